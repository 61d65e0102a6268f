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
t = buf.reshape(G, sh.n_layers + 1, 16).astype(np.int64)
names = ["norm+quant", "QKV gemv", "barrier", "attention", "barrier", "combine", "barrier", "load+WO gemv", "barrier",
         "norm+quant", "W13 gemv", "barrier", "quant h", "W2 gemv", "barrier"]
d = np.diff(t[:, : sh.n_layers, :], axis=2) / 1e3  # us, [G][L][15]
print(f"{shape} ctx {ctx}: grid {G}; per-layer phase time in us (mean over layers 2.., median/max over CTAs)")
for k, nm in enumerate(names):
    x = d[:, 2:, k].mean(axis=1)
    print(f"  {nm:14s} median {np.median(x):7.2f}  max {x.max():7.2f}  min {x.min():7.2f}")
layer = (t[:, 1:sh.n_layers, 0] - t[:, : sh.n_layers - 1, 0]) / 1e3
print(f"  layer total   median {np.median(layer[:, 2:].mean(axis=1)):.2f} us")
print(f"  cls: norm {np.median(t[:, -1, 1] - t[:, -1, 0]) / 1e3:.2f} us, gemv {np.median(t[:, -1, 2] - t[:, -1, 1]) / 1e3:.2f} us; whole kernel {(t[:, -1, 2].max() - t[:, 0, 0].min()) / 1e3:.1f} us")
