import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as e
pkg = e._pkg(); ql = pkg.QwenLib()
shape, ctx = "4b", 4096
path = pkg.checkpoint.ensure_checkpoint("/tmp/qwen3_b200_ckpt", shape, seed=1234, mode="fast")
gm = ql.open(path, ctx + 64)
for i in range(4): gm.forward_nocopy(7, ctx + i)
f = ql.lib.qwen_cuda_debug_tile_log; f.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
warp = int(sys.argv[1]) if len(sys.argv) > 1 else 0
n = f(gm.ctx, warp, None)
gm.forward_nocopy(7, ctx + 8)
buf = np.zeros(4 * n, np.uint64); f(gm.ctx, warp, buf.ctypes.data_as(C.c_void_p))
t = buf.reshape(4, n).astype(np.int64)
valid = np.nonzero(t[1] > 0)[0]
t0 = t[0][t[0] > 0].min()
print("tiles logged", len(valid), "first idx", valid[:5])
per_layer = len(valid) // 37
lo = 3 * per_layer
print(" it | issue(us)  wait_begin  wait_end  done | waited  busy  issue->ready")
for it in valid[lo: lo + per_layer + 4]:
    iss, wb, we, dn = [(t[k][it] - t0) / 1e3 for k in range(4)]
    print(f"{it:4d} | {iss:9.2f} {wb:9.2f} {we:9.2f} {dn:9.2f} | {we - wb:6.2f} {dn - we:6.2f} {we - iss:7.2f}")
