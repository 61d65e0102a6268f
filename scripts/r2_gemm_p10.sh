cd $GRAFT_REPO_ROOT
O=gpurun_out/r2_gemm_p10.log; : > $O
echo "== exact" >> $O
for k in 3 8; do QWEN_GEMM_PROF=$k timeout 120 python scripts/prefill_gemm_bench.py 2>&1 | grep "gemm prof\] C\|^4B\|^1.7" >> $O; done
echo "== fma fold" >> $O
for k in 3 8; do QWEN_GEMM_EXACT=0 QWEN_GEMM_PROF=$k timeout 120 python scripts/prefill_gemm_bench.py 2>&1 | grep "gemm prof\] C\|^4B\|^1.7" >> $O; done
cat $O
