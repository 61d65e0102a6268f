cd $GRAFT_REPO_ROOT
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:k_prefill_gemm_p --launch-skip 7 --launch-count 1 -o gpurun_out/r2_gemm_p_w13 -f python scripts/prefill_gemm_bench.py > gpurun_out/ncu_gemm.log 2>&1
tail -3 gpurun_out/ncu_gemm.log; ls -la gpurun_out/r2_gemm_p_w13.ncu-rep
