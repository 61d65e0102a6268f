"""Per-warp cycle split of the persistent kernel's GEMV phases (wait for the ring / dp4a loop / reduction + epilogue).
Needs a -DQW_UNITPROF build: make -C qwen3.c_b200/csrc OUT=../lib_prof EXTRA=-DQW_UNITPROF, then
QWEN3_LIB_PATH=qwen3.c_b200/lib_prof/libqwen3.so python scripts/unit_profile.py 4b 4096"""
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
L = sh.n_layers
up = buf[G * (L + 1) * 16:].reshape(G, 16, 5, 8).astype(np.float64)[:, :15]
print(f"{shape} ctx {ctx}: per layer and warp, cycles (mean over CTAs and warps; max over warps of the CTA mean)")
for mi, nm in enumerate(("QKV", "WO", "W13", "W2", "CLS")):
    div = L if mi < 4 else 1
    w, m, ep, units, bar, store, other = (up[:, :, mi, k] / div for k in range(7))
    tot = w + m + ep
    print(f"  {nm:4s} units/warp {units.mean():5.2f} (max {units.max():.0f})  load {w.mean():7.0f}  math {m.mean():7.0f}  epilogue {ep.mean():6.0f}  total {tot.mean():7.0f} (slowest warp {tot.max(axis=1).mean():7.0f})"
          f"  math/unit {m.sum() / max(units.sum(), 1):6.0f}  epi/unit {ep.sum() / max(units.sum(), 1):5.0f}"
          f"  | end barrier {bar.mean():6.0f}  store (thread 0) {store[:, 0].mean():6.0f}  loop overhead / issue {other.mean():6.0f}")
