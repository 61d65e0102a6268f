cd $GRAFT_REPO_ROOT
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:k_prefill_gemm_p --launch-skip 13 --launch-count 1 -o gpurun_out/r2_gemm_p_w2 -f python scripts/prefill_gemm_bench.py > gpurun_out/ncu_gemm2.log 2>&1
tail -2 gpurun_out/ncu_gemm2.log
