cd $GRAFT_REPO_ROOT
O=gpurun_out/r2_gemm_p14.log; : > $O
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "prefill_matmul_batch" 2>&1 | tail -2 >> $O
run() { echo "== $1" >> $O; timeout 120 python scripts/prefill_gemm_bench.py 2>&1 | grep "^4B\|^1.7" >> $O; }
run "auto"
QWEN_GEMM_N=192 run "N=192"
QWEN_GEMM_N=176 run "N=176"
QWEN_GEMM_N=144 run "N=144"
QWEN_GEMM_N=96 run "N=96"
cat $O
