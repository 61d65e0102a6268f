cd $GRAFT_REPO_ROOT
O=gpurun_out/r2_prefill_a1.log; : > $O
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "prefill" 2>&1 | tail -5 >> $O
for sh in 1.7b 4b; do
  echo "== $sh tiled attention + persistent GEMM" >> $O
  timeout 300 python scripts/prefill_once.py $sh 512 3 2>&1 | grep prefill >> $O
  echo "== $sh round-1 attention (QWEN_ATTN_V=1)" >> $O
  QWEN_ATTN_V=1 timeout 300 python scripts/prefill_once.py $sh 512 3 2>&1 | grep prefill >> $O
  echo "== $sh round-1 attention and GEMM" >> $O
  QWEN_ATTN_V=1 QWEN_GEMM_V=1 timeout 300 python scripts/prefill_once.py $sh 512 3 2>&1 | grep prefill >> $O
done
cat $O
