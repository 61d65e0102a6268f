"""One prefill of T tokens (used under ncu for the per-kernel launch list) -- prints device-side wall time."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as e
pkg = e._pkg(); ql = pkg.QwenLib()
shape = sys.argv[1] if len(sys.argv) > 1 else "1.7b"
T = int(sys.argv[2]) if len(sys.argv) > 2 else 512
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
path = pkg.checkpoint.ensure_checkpoint("/tmp/qwen3_b200_ckpt", shape, seed=1234, mode="fast")
gm = ql.open(path, T + 64)
toks = [int(t) for t in np.random.default_rng(0).integers(0, 1000, size=T)]
best = 1e9
for r in range(reps):
    t0 = time.perf_counter(); assert gm.prefill_nocopy(toks, 0); best = min(best, time.perf_counter() - t0)
print(f"{shape} prefill T={T}: {best*1e3:.2f} ms  {T/best:.0f} tok/s", flush=True)
gm.close()
