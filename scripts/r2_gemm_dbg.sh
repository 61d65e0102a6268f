cd $GRAFT_REPO_ROOT
O=gpurun_out/r2_gemm_dbg.log; : > $O
for d in 0 16 2 18 3 19 7 23 8 12 15 31; do
  echo "== QWEN_GEMM_DBG=$d (1 no tcgen05.ld, 2 no arithmetic, 4 no MMA, 8 no TMA, 16 no weight-scale load)" >> $O
  QWEN_GEMM_DBG=$d timeout 120 python scripts/prefill_gemm_bench.py 2>&1 | grep -v "^\[" | head -3 >> $O
done
cat $O
