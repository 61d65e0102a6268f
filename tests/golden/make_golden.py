"""Regenerates tests/golden/micro.bin and tests/golden/micro_golden.npz.

Runs ONLY in the build container (needs /root/reference):
  * micro.bin is written by the reference's OWN exporter (qwen3/weights.py:361 model_write,
    :137 quantize_q8_0) from a seeded reference `Transformer` container (qwen3/model.py:213);
  * every expected output in micro_golden.npz comes from the reference's OWN compiled C
    (oracle/_ref/libqwen3_ref_strict.so = src/{q8,model,forward}.c, gcc -O2 -DNDEBUG, serial).
Nothing from this repo's product or oracle restatement is used to produce expected values.

    python tests/golden/make_golden.py
"""
import contextlib
import ctypes as C
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import torch  # noqa: E402
from qwen3.model import ModelArgs, Transformer  # noqa: E402  (reference package)
from qwen3.weights import model_write  # noqa: E402

from oracle.binding import Q8Tensor, RefLib, build_ref, c_float_p, c_int8_p  # noqa: E402

MICRO = dict(dim=256, n_layers=2, n_heads=2, n_kv_heads=1, head_dim=128, vocab_size=256,
             hidden_dim=256, max_seq_len=128)
TOKENS = [7, 200, 31, 31, 0, 255, 128, 64, 99, 1, 2, 3]


def main():
    assert build_ref(), "needs /root/reference"
    torch.manual_seed(20261018)
    model = Transformer(ModelArgs(**MICRO))
    with torch.no_grad():
        # healthy spreads so logits separate: N(0, 1/sqrt(fan_in)) everywhere
        for name, p in model.named_parameters():
            if p.ndim == 2:
                p.normal_(0.0, 1.0 / np.sqrt(p.shape[1]))
            else:
                p.copy_(1.0 + 0.1 * torch.randn_like(p))
        model.tok_embeddings.weight.mul_(4.0)  # tied with output: logit std ~ 4
    path = os.path.join(HERE, "micro.bin")
    with contextlib.redirect_stdout(io.StringIO()):
        model_write(model, path, 64)

    ref = RefLib("strict")
    lib = ref.lib
    rng = np.random.default_rng(99)
    out = {"tokens": np.array(TOKENS, np.int32)}

    # ---- whole forward: logits at every position, final KV cache -------------
    m = ref.open(path, 0)
    p = m.contents.params
    logits = np.stack([ref.forward(m, t, pos) for pos, t in enumerate(TOKENS)])
    n_kv = p.n_layers * p.seq_len * p.n_kv_heads * p.head_dim
    kc = np.ctypeslib.as_array(m.contents.state.k_cache, shape=(n_kv,)).copy()
    vc = np.ctypeslib.as_array(m.contents.state.v_cache, shape=(n_kv,)).copy()
    shape = (p.n_layers, p.seq_len, p.n_kv_heads * p.head_dim)
    out["logits"] = logits
    out["k_cache"] = kc.reshape(shape)[:, : len(TOKENS)].copy()
    out["v_cache"] = vc.reshape(shape)[:, : len(TOKENS)].copy()

    # ---- attention() on that cache, fresh q ----------------------------------
    P = p.n_heads * p.head_dim
    q = rng.standard_normal(P).astype(np.float32)
    C.memmove(m.contents.state.q, q.ctypes.data, 4 * P)
    lib.attention(m, 1, len(TOKENS) - 1)
    out["att_q"] = q
    out["att_out"] = np.ctypeslib.as_array(m.contents.state.x_rms_norm, shape=(P,)).copy()
    ref.close(m)

    # ---- q8_quantize: ordinary, zero group, ties at .5, clamp, tiny/huge -----
    x = rng.standard_normal(64 * 12).astype(np.float32)
    x[64:128] = 0.0                               # all-zero group -> scale 1e-6
    x[128:192] = np.float32(127.0) * np.linspace(-1, 1, 64, dtype=np.float32)  # exact codes
    x[192:256] = (np.arange(64) - 31.5).astype(np.float32) * 2  # halves after /scale
    x[256:320] *= np.float32(1e-30)
    x[320:384] *= np.float32(1e30)
    x[384] = np.float32(5e4)                      # one outlier squeezes the rest to ~0
    q8 = np.zeros(x.size, np.int8)
    s8 = np.zeros(x.size // 64, np.float32)
    t = Q8Tensor(s8.ctypes.data_as(c_float_p), q8.ctypes.data_as(c_int8_p))
    xin = x.copy()
    lib.q8_quantize(C.byref(t), xin.ctypes.data_as(c_float_p), x.size, 64)
    out["quant_x"], out["quant_q"], out["quant_s"] = x, q8, s8
    deq = np.zeros(x.size, np.float32)
    lib.q8_dequantize(C.byref(t), deq.ctypes.data_as(c_float_p), x.size, 64)
    out["dequant_x"] = deq

    # ---- matmul: d x n with n = 5 groups, extreme codes in one row ------------
    n, d = 320, 37
    wq = rng.integers(-127, 128, size=d * n, dtype=np.int8)
    wq[:n] = 127
    wq[n:2 * n] = -127
    ws = rng.uniform(1e-3, 2e-2, size=d * n // 64).astype(np.float32)
    xa = rng.standard_normal(n).astype(np.float32) * 3
    xa[:64] = 1e3
    xq = np.zeros(n, np.int8)
    xs = np.zeros(n // 64, np.float32)
    tx = Q8Tensor(xs.ctypes.data_as(c_float_p), xq.ctypes.data_as(c_int8_p))
    lib.q8_quantize(C.byref(tx), xa.ctypes.data_as(c_float_p), n, 64)
    xq[:64] = 127
    tw = Q8Tensor(ws.ctypes.data_as(c_float_p), wq.ctypes.data_as(c_int8_p))
    mo = np.zeros(d, np.float32)
    lib.matmul(mo.ctypes.data_as(c_float_p), C.byref(tx), C.byref(tw), n, d, 64)
    out.update(mm_xq=xq, mm_xs=xs, mm_wq=wq, mm_ws=ws, mm_out=mo, mm_nd=np.array([n, d], np.int32))

    # ---- rmsnorm (D-sized and head-sized), softmax, rotary, swiglu ------------
    for tag, size in (("big", 2560), ("head", 128)):
        xv = (rng.standard_normal(size) * 5).astype(np.float32)
        wv = (1 + 0.1 * rng.standard_normal(size)).astype(np.float32)
        ov = np.zeros(size, np.float32)
        lib.rmsnorm(ov.ctypes.data_as(c_float_p), xv.ctypes.data_as(c_float_p), wv.ctypes.data_as(c_float_p), size)
        out[f"rms_{tag}_x"], out[f"rms_{tag}_w"], out[f"rms_{tag}_out"] = xv, wv, ov
    sm = (rng.standard_normal(777) * 6).astype(np.float32)
    out["softmax_x"] = sm.copy()
    lib.softmax(sm.ctypes.data_as(c_float_p), sm.size)
    out["softmax_out"] = sm
    rot_in = rng.standard_normal(128).astype(np.float32)
    out["rot_x"] = rot_in
    for pos in (0, 1, 777, 4095, 32767):
        r = rot_in.copy()
        lib.rotary(r.ctypes.data_as(c_float_p), 128, pos)
        out[f"rot_out_{pos}"] = r
    a = (rng.standard_normal(1000) * 4).astype(np.float32)
    a[:4] = [0.0, -100.0, 100.0, -0.0]
    b = rng.standard_normal(1000).astype(np.float32)
    out["swiglu_x1"], out["swiglu_x3"] = a.copy(), b
    lib.swiglu(a.ctypes.data_as(c_float_p), b.ctypes.data_as(c_float_p), a.size)
    out["swiglu_out"] = a
    out["silu_x"] = np.array([-20, -1, -0.5, 0, 0.5, 1, 20], np.float32)
    out["silu_out"] = np.array([lib.silu(float(v)) for v in out["silu_x"]], np.float32)
    out["sigmoid_out"] = np.array([lib.sigmoid(float(v)) for v in out["silu_x"]], np.float32)

    np.savez_compressed(os.path.join(HERE, "micro_golden.npz"), **out)
    print("wrote", path, os.path.getsize(path), "bytes and micro_golden.npz",
          os.path.getsize(os.path.join(HERE, "micro_golden.npz")), "bytes")
    print("logit std", logits.std(), "argmax", logits.argmax(1))


if __name__ == "__main__":
    main()
