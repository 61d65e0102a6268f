cd $GRAFT_REPO_ROOT
O=gpurun_out/r2_prefill_a4.log; : > $O
timeout 600 python scripts/prefill_once.py 4b 512 3 2>&1 | grep prefill >> $O
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_prefill_4b_launches.csv python scripts/prefill_once.py 4b 512 1 > gpurun_out/ncu_prefill.log 2>&1
tail -3 gpurun_out/ncu_prefill.log >> $O
cat $O
