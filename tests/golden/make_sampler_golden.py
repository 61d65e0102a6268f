"""Writes tests/golden/sampler_golden.npz: outputs of the REFERENCE's own sampler.c / xorshift.c
(oracle/_ref/libqwen3_ref_sampler.so, compiled unchanged by oracle/Makefile) on seeded logits.
Run in the build container (where /root/reference exists): python tests/golden/make_sampler_golden.py"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.binding import RefSampler  # noqa: E402


def cases():
    """(vocab, temperature, top_p, seed, logit scale): what the CLI can produce -- greedy-like (t -> 1e-6), defaults, flat."""
    out = []
    for i, (V, t, p) in enumerate([(512, 1.0, 0.9), (512, 0.0, 0.9), (2048, 0.7, 0.8), (2048, 1.5, 0.95), (4096, 1.0, 0.5),
                                   (4096, 0.3, 0.99), (1000, 1.0, 1.0), (1000, 2.0, 0.0), (151936, 0.7, 0.8), (151936, 1.0, 0.9)]):
        out.append((V, t, p, 1000 + i, 4.0 if i % 2 else 8.0))
    return out


def main():
    ref = RefSampler()
    rec = {}
    for ci, (V, t, p, seed, scale) in enumerate(cases()):
        rng = np.random.default_rng(seed)
        s = ref.create(V, t, p, seed)
        toks = []
        logits_all = []
        for step in range(6):
            lg = (rng.standard_normal(V) * scale).astype(np.float32)
            logits_all.append(lg)
            toks.append(ref.sample(s, lg))
        rec[f"c{ci}_meta"] = np.array([V, seed], np.int64)
        rec[f"c{ci}_tp"] = np.array([t, p, scale], np.float32)
        rec[f"c{ci}_clamped"] = np.array([s.contents.temperature, s.contents.top_p], np.float32)
        rec[f"c{ci}_tokens"] = np.array(toks, np.int32)
        rec[f"c{ci}_seed_after"] = np.array([s.contents.seed], np.uint64)
        if V <= 4096:
            rec[f"c{ci}_logits"] = np.stack(logits_all)
        ref.free(s)
    st = C.c_uint64(42)
    rec["xorshift42"] = np.array([ref.lib.xorshift_float(C.byref(st)) for _ in range(64)], np.float32)
    rec["xorshift42_state"] = np.array([st.value], np.uint64)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "sampler_golden.npz"), **rec)
    print("wrote sampler_golden.npz with", len(cases()), "cases")


if __name__ == "__main__":
    main()
