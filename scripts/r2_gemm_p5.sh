cd $GRAFT_REPO_ROOT
O=gpurun_out/r2_gemm_p5.log; : > $O
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "prefill_matmul_batch" 2>&1 | tail -3 >> $O
run() { echo "== $1" >> $O; timeout 120 python scripts/prefill_gemm_bench.py 2>&1 | grep -v "^\[" | head -${2:-4} >> $O; }
run "persistent (auto N)"
QWEN_GEMM_N=160 run "N=160"
QWEN_GEMM_N=128 run "N=128"
QWEN_GEMM_N=96 run "N=96"
QWEN_GEMM_N=80 run "N=80"
QWEN_GEMM_N=64 run "N=64"
QWEN_GEMM_V=1 run "round-1 kernel"
cat $O
