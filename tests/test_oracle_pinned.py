"""Pins the CPU restatement (oracle/qwen3_oracle.c) to the reference.

(a) against tests/golden/ -- outputs of the reference's own compiled C on a checkpoint written
    by the reference's own exporter (tests/golden/make_golden.py);
(b) against oracle/_ref/libqwen3_ref_strict.so on fresh random inputs, when that library
    travelled with the repo.
Everything is required to be BIT-IDENTICAL: same operations in the same order in fp32.
"""
import ctypes as C
import os

import numpy as np
import pytest

from oracle.binding import Q8Tensor, RefLib, c_float_p, c_int8_p

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(os.path.join(GOLD, "micro_golden.npz")))


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape and a.dtype == b.dtype
    assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), f"max abs diff {np.abs(a - b).max()}"


def test_golden_forward_logits_and_kv(oracle, gold):
    with oracle.open(os.path.join(GOLD, "micro.bin")) as m:
        for pos, t in enumerate(gold["tokens"]):
            same(m.forward(int(t), pos), gold["logits"][pos])
        k, v = m.kv()
        n = len(gold["tokens"])
        same(k[:, :n], gold["k_cache"])
        same(v[:, :n], gold["v_cache"])


def test_golden_attention(oracle, gold):
    L, T, kvd = gold["k_cache"].shape
    k = np.zeros((T, kvd), np.float32)
    k[:] = gold["k_cache"][1]
    v = gold["v_cache"][1].copy()
    out = oracle.attention(gold["att_q"], k, v, n_heads=2, n_kv_heads=1, head_dim=128, seq_len=T, pos=T - 1)
    same(out, gold["att_out"])


def test_golden_quantize_edges(oracle, gold):
    q, s = oracle.q8_quantize(gold["quant_x"])
    same(q, gold["quant_q"])
    same(s, gold["quant_s"])
    assert s[1] == np.float32(1e-6) and not q[64:128].any()      # all-zero group
    assert q.min() >= -127 and q.max() <= 127
    same(oracle.q8_dequantize(q, s), gold["dequant_x"])


def test_golden_matmul(oracle, gold):
    n, d = map(int, gold["mm_nd"])
    out = oracle.matmul(gold["mm_xq"], gold["mm_xs"], gold["mm_wq"], gold["mm_ws"], n, d)
    same(out, gold["mm_out"])
    dots = oracle.group_dots(gold["mm_xq"], gold["mm_wq"], n, d)
    assert dots[0, 0] == 64 * 127 * 127 and dots[1, 0] == -64 * 127 * 127  # extreme group, exact
    # the fp32 fold is reproducible from the integer dots alone
    acc = np.zeros(d, np.float32)
    ws = gold["mm_ws"].reshape(d, -1)
    for g in range(n // 64):
        acc = acc + (dots[:, g].astype(np.float32) * ws[:, g]) * gold["mm_xs"][g]
    same(acc, gold["mm_out"])


def test_golden_elementwise(oracle, gold):
    for tag in ("big", "head"):
        same(oracle.rmsnorm(gold[f"rms_{tag}_x"], gold[f"rms_{tag}_w"]), gold[f"rms_{tag}_out"])
    same(oracle.softmax(gold["softmax_x"]), gold["softmax_out"])
    for pos in (0, 1, 777, 4095, 32767):
        same(oracle.rotary(gold["rot_x"], 128, pos), gold[f"rot_out_{pos}"])
    same(oracle.swiglu(gold["swiglu_x1"], gold["swiglu_x3"]), gold["swiglu_out"])
    silu = np.array([oracle.lib.orc_silu(float(v)) for v in gold["silu_x"]], np.float32)
    sig = np.array([oracle.lib.orc_sigmoid(float(v)) for v in gold["silu_x"]], np.float32)
    same(silu, gold["silu_out"])
    same(sig, gold["sigmoid_out"])


def test_micro_header_matches_reference_exporter(pkg):
    """The golden checkpoint was written by the reference exporter; our Shape arithmetic must
    predict its size, i.e. our reading of the layout (checkpoint.py) is the exporter's."""
    import struct
    with open(os.path.join(GOLD, "micro.bin"), "rb") as f:
        hdr = struct.unpack("<12i", f.read(48))
    assert hdr[0] == pkg.checkpoint.MAGIC and hdr[1] == 1
    sh = pkg.checkpoint.Shape("micro", dim=hdr[2], hidden_dim=hdr[3], n_layers=hdr[4], n_heads=hdr[5],
                              n_kv_heads=hdr[6], vocab_size=hdr[7], seq_len=hdr[8], head_dim=hdr[9],
                              shared_classifier=hdr[10], group_size=hdr[11])
    assert sh.file_bytes() == os.path.getsize(os.path.join(GOLD, "micro.bin"))


# --------------------------------------------------------------------------- (b)
needs_ref = pytest.mark.skipif(not RefLib.available("strict"), reason="oracle/_ref not built here")


@needs_ref
@pytest.mark.parametrize("shape", ["tiny", "tiny-untied"])
def test_forward_bit_identical_to_compiled_reference(oracle, pkg, ckpt_dir, shape):
    path = pkg.checkpoint.ensure_checkpoint(ckpt_dir, shape, seed=7)
    ref = RefLib("strict")
    rm = ref.open(path, 48)
    rng = np.random.default_rng(3)
    V = pkg.checkpoint.SHAPES[shape].vocab_size
    with oracle.open(path, 48) as om:
        for pos in range(20):
            t = int(rng.integers(0, V))
            same(om.forward(t, pos), ref.forward(rm, t, pos))
    ref.close(rm)


@needs_ref
def test_ops_bit_identical_to_compiled_reference(oracle):
    lib = RefLib("strict").lib
    rng = np.random.default_rng(11)
    for n in (64, 1024, 2560, 9728):
        x = (rng.standard_normal(n) * rng.uniform(0.01, 50)).astype(np.float32)
        q, s = np.zeros(n, np.int8), np.zeros(n // 64, np.float32)
        t = Q8Tensor(s.ctypes.data_as(c_float_p), q.ctypes.data_as(c_int8_p))
        lib.q8_quantize(C.byref(t), x.copy().ctypes.data_as(c_float_p), n, 64)
        oq, os_ = oracle.q8_quantize(x)
        same(oq, q)
        same(os_, s)
        d = 19
        wq = rng.integers(-127, 128, size=d * n, dtype=np.int8)
        ws = rng.uniform(1e-4, 1e-2, size=d * n // 64).astype(np.float32)
        tw = Q8Tensor(ws.ctypes.data_as(c_float_p), wq.ctypes.data_as(c_int8_p))
        out = np.zeros(d, np.float32)
        lib.matmul(out.ctypes.data_as(c_float_p), C.byref(t), C.byref(tw), n, d, 64)
        same(oracle.matmul(q, s, wq, ws, n, d), out)
        w = rng.standard_normal(n).astype(np.float32)
        ro = np.zeros(n, np.float32)
        lib.rmsnorm(ro.ctypes.data_as(c_float_p), x.copy().ctypes.data_as(c_float_p), w.ctypes.data_as(c_float_p), n)
        same(oracle.rmsnorm(x, w), ro)


@needs_ref
def test_numpy_weight_quantiser_matches_reference_exporter(pkg):
    import sys
    if not os.path.isdir("/root/reference/qwen3"):
        pytest.skip("reference python package not present")
    sys.path.insert(0, "/root/reference")
    import torch
    from qwen3.weights import quantize_q8_0
    w = torch.randn(64 * 300) * 0.02
    r = quantize_q8_0(w, 64)
    q, s = pkg.checkpoint.quantize_q8_0(w.numpy(), 64)
    same(q, r.quant.reshape(-1).numpy())
    same(s, r.scale.numpy())
