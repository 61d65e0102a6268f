cd $GRAFT_REPO_ROOT
O=gpurun_out/r2_gemm_p3.log; : > $O
echo "== dbg 15 prof" >> $O
QWEN_GEMM_PROF=1 QWEN_GEMM_DBG=15 timeout 120 python scripts/prefill_gemm_bench.py 2>&1 | grep "gemm prof\|4B qkv" >> $O
echo "== dbg 0 prof" >> $O
QWEN_GEMM_PROF=1 timeout 120 python scripts/prefill_gemm_bench.py 2>&1 | grep "gemm prof\|4B qkv" >> $O
cat $O
