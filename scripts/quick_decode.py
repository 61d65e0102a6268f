"""Quick A/B timing of the persistent decode kernel: `QWEN3_LIB_PATH=<variant .so> python scripts/quick_decode.py 4b 4096 64`
prints ms/token (CUDA events, qwen_cuda_time_decode). Variants are built side by side with `make -C qwen3.c_b200/csrc OUT=../lib_x`."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as e
pkg = e._pkg(); ql = pkg.QwenLib()
shape = sys.argv[1] if len(sys.argv) > 1 else "4b"
ctx = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 64
path = pkg.checkpoint.ensure_checkpoint("/tmp/qwen3_b200_ckpt", shape, seed=1234, mode="fast")
gm = ql.open(path, ctx + 2 * steps + 64)
best = 1e9
for rep in range(2):
    ms, _ = gm.time_decode(7, ctx, steps, 4)
    best = min(best, ms / steps)
print(f"{os.environ.get('QWEN3_LIB_PATH', 'default')} {shape} ctx {ctx}: {best:.4f} ms/token = {1e3 / best:.1f} tok/s", flush=True)
gm.close()
