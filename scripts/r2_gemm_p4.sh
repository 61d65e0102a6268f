cd $GRAFT_REPO_ROOT
O=gpurun_out/r2_gemm_p4.log; : > $O
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "prefill_matmul_batch" 2>&1 | tail -3 >> $O
run() { echo "== $1" >> $O; timeout 120 python scripts/prefill_gemm_bench.py 2>&1 | grep -v "^\[" | head -${2:-4} >> $O; }
run "persistent"
QWEN_GEMM_DBG=2 run "dbg 2: no math" 3
QWEN_GEMM_DBG=3 run "dbg 3: no ld no math" 3
QWEN_GEMM_DBG=15 run "dbg 15: skeleton" 3
echo "== prof" >> $O
QWEN_GEMM_PROF=1 timeout 120 python scripts/prefill_gemm_bench.py 2>&1 | grep "gemm prof" | head -14 >> $O
cat $O
