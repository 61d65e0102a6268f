cd $GRAFT_REPO_ROOT
scripts/ubench/hop > gpurun_out/r2_hop.log 2>&1
timeout 300 python scripts/prefill_gemm_bench.py > gpurun_out/r2_gemm_bench.log 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "prefill" 2>&1 | tail -4 > gpurun_out/r2_prefill_tests.log
timeout 200 python scripts/prefill_once.py 4b 512 > gpurun_out/r2_prefill_4b.log 2>&1
