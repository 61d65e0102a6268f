"""A few decode steps of one workload (used under ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as e
pkg = e._pkg(); ql = pkg.QwenLib()
shape = sys.argv[1] if len(sys.argv) > 1 else "4b"
ctx = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
path = pkg.checkpoint.ensure_checkpoint("/tmp/qwen3_b200_ckpt", shape, seed=1234, mode="fast")
gm = ql.open(path, ctx + 64)
for i in range(steps):
    assert gm.forward_nocopy(7, ctx + i)
gm.close()
print("ok")
