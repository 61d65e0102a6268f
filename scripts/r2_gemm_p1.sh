cd $GRAFT_REPO_ROOT
O=gpurun_out/r2_gemm_p1.log; : > $O
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "prefill_matmul_batch" 2>&1 | tail -15 >> $O
echo "== persistent kernel" >> $O
timeout 120 python scripts/prefill_gemm_bench.py 2>&1 | grep -v "^\[" >> $O
echo "== round-1 kernel (QWEN_GEMM_V=1)" >> $O
QWEN_GEMM_V=1 timeout 120 python scripts/prefill_gemm_bench.py 2>&1 | grep -v "^\[" >> $O
for d in 2 3 7 31; do
  echo "== persistent, QWEN_GEMM_DBG=$d" >> $O
  QWEN_GEMM_DBG=$d timeout 120 python scripts/prefill_gemm_bench.py 2>&1 | grep -v "^\[" | head -3 >> $O
done
cat $O
