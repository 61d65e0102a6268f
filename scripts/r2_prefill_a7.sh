cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "prefill" 2>&1 | grep -v "^\[" | grep -B5 -A25 "Error\|assert" | head -80 > gpurun_out/r2_prefill_a7.log
cat gpurun_out/r2_prefill_a7.log
