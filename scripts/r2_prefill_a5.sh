cd $GRAFT_REPO_ROOT
O=gpurun_out/r2_prefill_a5.log; : > $O
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "prefill" 2>&1 | tail -3 >> $O
for sh in 1.7b 4b; do
  timeout 300 python scripts/prefill_once.py $sh 512 3 2>&1 | grep prefill >> $O
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_prefill_4b_launches.csv python scripts/prefill_once.py 4b 512 1 > gpurun_out/ncu_prefill.log 2>&1
cat $O
