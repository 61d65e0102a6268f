cd $GRAFT_REPO_ROOT
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:k_attn_prefill_t --launch-skip 20 --launch-count 1 -o gpurun_out/r2_attn_t -f python scripts/prefill_once.py 4b 512 1 > gpurun_out/ncu_attn.log 2>&1
tail -2 gpurun_out/ncu_attn.log
