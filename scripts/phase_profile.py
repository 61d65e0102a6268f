"""Per-phase timing of the persistent decode kernel from in-kernel globaltimer stamps."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as e
pkg = e._pkg(); ql = pkg.QwenLib()
shape = sys.argv[1] if len(sys.argv) > 1 else "4b"
ctx = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
path = pkg.checkpoint.ensure_checkpoint("/tmp/qwen3_b200_ckpt", shape, seed=1234, mode="fast")
sh = pkg.checkpoint.SHAPES[shape]
gm = ql.open(path, ctx + 64)
for i in range(4): gm.forward_nocopy(7, ctx + i)
ql.lib.qwen_cuda_debug_profile_enable.argtypes = [C.c_void_p]
ql.lib.qwen_cuda_debug_profile_read.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
n = ql.lib.qwen_cuda_debug_profile_enable(gm.ctx)
gm.forward_nocopy(7, ctx + 8)
buf = np.zeros(n, np.uint64)
G = ql.lib.qwen_cuda_debug_profile_read(gm.ctx, buf.ctypes.data_as(C.c_void_p), n)
t = buf[: G * (sh.n_layers + 1) * 16].reshape(G, sh.n_layers + 1, 16).astype(np.int64)
# stamp slots per layer: 4k = phase k start, 4k+2 = after its prologue, 4k+3 = after its GEMV; 5 = after attention + combine
names = [("poll x + norm + quant", 0, 2), ("QKV gemv", 2, 3), ("attn: poll q, norm, rope", 4, 1), ("attn: tiles", 1, 9), ("attn: merge + publish", 9, 13), ("combine", 13, 5), ("poll att_q", 5, 6), ("WO gemv", 6, 7),
         ("poll x + norm + quant", 8, 10), ("W13 gemv", 10, 11), ("poll h + quant", 12, 14), ("W2 gemv", 14, 15)]
print(f"{shape} ctx {ctx}: grid {G}; per-layer phase time in us (mean over layers 2.., median/max over CTAs)")
L = sh.n_layers
for nm, a, b in names:
    x = ((t[:, 2:L, b] - t[:, 2:L, a]) / 1e3).mean(axis=1)
    print(f"  {nm:26s} median {np.median(x):7.2f}  max {x.max():7.2f}  min {x.min():7.2f}")
layer = (t[:, 1:L, 0] - t[:, : L - 1, 0]) / 1e3
print(f"  layer total   median {np.median(layer[:, 2:].mean(axis=1)):.2f} us")
for k, nm in ((3, "QKV done"), (5, "attention done"), (7, "WO done"), (11, "W13 done"), (15, "W2 done")):
    f = t[:, 2:L, k]
    sk = (f - f.min(axis=0, keepdims=True)).mean(axis=1) / 1e3
    print(f"  skew {nm:15s} median {np.median(sk):5.2f}  p90 {np.percentile(sk, 90):5.2f}  max {sk.max():5.2f} us")
print(f"  cls: prologue {np.median(t[:, -1, 2] - t[:, -1, 0]) / 1e3:.2f} us, gemv {np.median(t[:, -1, 3] - t[:, -1, 2]) / 1e3:.2f} us; whole kernel {(t[:, -1, 3].max() - t[:, 0, 0].min()) / 1e3:.1f} us")
