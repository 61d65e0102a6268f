cd $GRAFT_REPO_ROOT
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_prefill_17b_launches.csv python scripts/prefill_once.py 1.7b 512 1 > gpurun_out/ncu_prefill17.log 2>&1
tail -2 gpurun_out/ncu_prefill17.log
