cd $GRAFT_REPO_ROOT
O=gpurun_out/r2_gemm_p2.log; : > $O
run() { echo "== $1" >> $O; timeout 120 python scripts/prefill_gemm_bench.py 2>&1 | grep -v "^\[" | head -3 >> $O; }
QWEN_GEMM_DBG=15 run "dbg 15: no TMA, no MMA, no ld, no math"
QWEN_GEMM_DBG=39 run "dbg 39: no scale box, no MMA, no ld, no math"
QWEN_GEMM_DBG=32 run "dbg 32: no scale box"
QWEN_GEMM_DBG=8 run "dbg 8: no TMA"
QWEN_GEMM_N=128 run "N=128"
QWEN_GEMM_N=128 QWEN_GEMM_DBG=15 run "N=128 dbg 15"
QWEN_GEMM_N=64 run "N=64"
cat $O
