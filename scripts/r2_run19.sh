cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests/test_gpu_strict.py -x -q -s 2>&1 | tail -40 > gpurun_out/r2_strict.log
timeout 900 python -m pytest tests/test_gpu_parity.py -q 2>&1 | tail -15 > gpurun_out/r2_parity_full.log
