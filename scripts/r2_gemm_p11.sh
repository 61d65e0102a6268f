cd $GRAFT_REPO_ROOT
O=gpurun_out/r2_gemm_p11.log; : > $O
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "prefill_matmul_batch" 2>&1 | tail -2 >> $O
for k in 3 8 13; do QWEN_GEMM_PROF=$k timeout 120 python scripts/prefill_gemm_bench.py 2>&1 | grep "gemm prof\] C\|^4B\|^1.7" >> $O; done
QWEN_GEMM_PROF=11 timeout 120 python scripts/prefill_gemm_bench.py 2>&1 | grep "gemm prof\] g0[1-6]" >> $O
cat $O
