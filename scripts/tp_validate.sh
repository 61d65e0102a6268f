#!/bin/bash
# Multi-GPU validation of the tensor-parallel persistent kernel: parity worker on N ranks, then the TP bench.
# usage: scripts/tp_validate.sh N [shapes...]   (run under gpurun --gpus N)
N=${1:-2}; shift
SHAPES=${@:-"8b-l2 32b-l2"}
mkdir -p gpurun_out /tmp/tpck
PATHS=$(python - <<PY
import sys
sys.path.insert(0, ".")
import __graft_entry__ as e
pkg = e._pkg()
print(" ".join(pkg.checkpoint.ensure_checkpoint("/tmp/tpck", n, seed=11) for n in "$SHAPES".split()))
PY
)
echo "checkpoints: $PATHS"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tests/tp_gpu_worker.py $PATHS > gpurun_out/tp${N}_worker.log 2>&1; cp gpurun_out/tp${N}_worker.log gpurun_out/r2_tp${N}_worker.log
echo "worker rc=$?"; grep -E "TP_GPU_OK|Error|assert" gpurun_out/tp${N}_worker.log | head -5
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 128 --warmup 8 > gpurun_out/bench_tp${N}.json 2> gpurun_out/bench_tp${N}.err
echo "bench rc=$?"; cat gpurun_out/bench_tp${N}.json; tail -2 gpurun_out/bench_tp${N}.err
