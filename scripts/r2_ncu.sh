cd $GRAFT_REPO_ROOT
python scripts/one_step.py 4b 4096 6 > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python scripts/one_step.py 4b 4096 6 > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_decode_pw -s 4 -c 1 -o gpurun_out/r2_k_decode_pw python scripts/one_step.py 4b 4096 6 > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"; ls -la gpurun_out/r2_k_decode_pw.ncu-rep; tail -3 gpurun_out/ncu_full.log
