cd $GRAFT_REPO_ROOT
O=gpurun_out/r2_gemm_p12.log; : > $O
for k in 3 8 13; do QWEN_GEMM_PROF=$k timeout 120 python scripts/prefill_gemm_bench.py 2>&1 | grep "gemm prof\] [Cp]" >> $O; done
cat $O
