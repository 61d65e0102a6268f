#!/bin/bash
# One 8-GPU call: tensor-parallel parity workers at TP = 8 and TP = 4 on the real 8B / 32B layer shapes (green logs go to
# profiles/), then BASELINE config 5: 32B shape, TP = 8, 2048-token prefill + decode near the 32768-position cap.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out /tmp/tpck
df -h /tmp | tail -1; free -g | head -2
PATHS=$(python - <<PY
import sys
sys.path.insert(0, ".")
import __graft_entry__ as e
pkg = e._pkg()
print(" ".join(pkg.checkpoint.ensure_checkpoint("/tmp/tpck", n, seed=11) for n in "8b-l2 32b-l2".split()))
PY
)
for N in 8 4; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N tests/tp_gpu_worker.py $PATHS > gpurun_out/r2_tp${N}_worker.log 2>&1
  echo "tp$N worker rc=$?"; grep -E "TP_GPU_OK|Error|assert|\[tp" gpurun_out/r2_tp${N}_worker.log | head -6
done
FREE=$(free -g | awk '/Mem:/{print $7}')
if [ "$FREE" -lt 100 ]; then echo "only $FREE GB of host memory available: config 5 skipped"; exit 0; fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29550 bench.py --gpus 8 --steps 64 --warmup 4 --workload 32b-decode-ctx32k --no-cpu-baseline --no-tp1 > gpurun_out/r2_bench_config5.json 2> gpurun_out/r2_bench_config5.err
echo "config5 rc=$?"; cat gpurun_out/r2_bench_config5.json; grep -v "^\[Params\]" gpurun_out/r2_bench_config5.err | tail -5
