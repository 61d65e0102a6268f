cd $GRAFT_REPO_ROOT
O=gpurun_out/r2_gemm_p8.log; : > $O
QWEN_GEMM_PROF=11 timeout 120 python scripts/prefill_gemm_bench.py 2>&1 | grep "gemm prof" | head -14 >> $O
cat $O
