cd $GRAFT_REPO_ROOT
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | grep -v "^\[" | tail -6 > gpurun_out/r2_final_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_final_smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err
timeout 600 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/r2_final_bench_ref.json 2> gpurun_out/r2_final_bench_ref.err
timeout 300 python scripts/prefill_once.py 1.7b 512 3 2>&1 | grep prefill > gpurun_out/r2_final_prefill.log
timeout 300 python scripts/prefill_once.py 4b 512 3 2>&1 | grep prefill >> gpurun_out/r2_final_prefill.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_prefill_17b_launches.csv python scripts/prefill_once.py 1.7b 512 1 > /dev/null 2>&1
tail -3 gpurun_out/r2_final_tests.log; tail -2 gpurun_out/r2_final_smoke.log; cut -c1-400 gpurun_out/r2_final_bench.json; cut -c1-300 gpurun_out/r2_final_bench_ref.json; cat gpurun_out/r2_final_prefill.log
