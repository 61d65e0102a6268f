import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as e
pkg = e._pkg(); ql = pkg.QwenLib()
shape, ctx = "4b", 4096
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
L = sh.n_layers
# arrival times at each barrier relative to the earliest arrival, per layer
names = {2: "after QKV", 4: "after attn", 6: "after combine", 8: "after WO", 11: "after W13", 14: "after W2"}
for k, nm in names.items():
    arr = t[:, 2:L, k]                      # [G][layers] arrival stamp
    rel = (arr - arr.min(axis=0)) / 1e3     # us after the first arriver
    lastcta = rel.argmax(axis=0)
    print(f"{nm:14s}: spread median {np.median(rel.max(axis=0)):.2f} us; mean lateness per CTA: top5 {np.sort(rel.mean(axis=1))[-5:].round(2)} CTAs {np.argsort(rel.mean(axis=1))[-5:]}; "
          f"corr of lateness between consecutive layers {np.corrcoef(rel[:, 3], rel[:, 4])[0,1]:.2f}")
# exit time from barrier relative to last arrival
for k, nm in names.items():
    arr = t[:, 2:L, k]; out = t[:, 2:L, k + 1]
    lat = (out - arr.max(axis=0)) / 1e3
    print(f"{nm:14s}: release latency after the last arrival: median {np.median(lat):.2f} us, max {lat.max():.2f}")
np.set_printoptions(linewidth=250, precision=1, suppress=True)
for k, nm in names.items():
    arr = t[:, 2:L, k]; rel = ((arr - arr.min(axis=0)) / 1e3).mean(axis=1)
    print(nm, "lateness per CTA (us):"); print(rel.reshape(-1, 37))
