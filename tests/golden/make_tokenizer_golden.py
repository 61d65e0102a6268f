"""Writes tests/golden/tokenizer_golden.npz: the ids the REFERENCE's own tokenizer.c (compiled into
oracle/_ref/libqwen3_ref_tokenizer.so by oracle/Makefile) produces for the texts of tests/test_tokenizer_cpu.py on the
synthetic vocabulary that test writes. Run here (where /root/reference exists); the fixture travels to the GPU box."""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import test_tokenizer_cpu as T  # noqa: E402

with tempfile.TemporaryDirectory() as d:
    prefix = os.path.join(d, "vocab")
    T.write_vocab(prefix)
    ref = T.Tok(T.REF, prefix)
    cases = T.texts(np.random.default_rng(3))
    np.savez(T.GOLD, **{f"ids{i}": np.array(ref.encode(t), np.int32) for i, t in enumerate(cases)})
    ref.close()
print("wrote", T.GOLD)
