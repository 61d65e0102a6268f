cd $GRAFT_REPO_ROOT
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | grep -v "^\[" | tail -15 > gpurun_out/r2_full1_tests.log
timeout 900 python bench.py > gpurun_out/r2_full1_bench.json 2> gpurun_out/r2_full1_bench.err
tail -5 gpurun_out/r2_full1_tests.log; cat gpurun_out/r2_full1_bench.json | cut -c1-3000; tail -5 gpurun_out/r2_full1_bench.err
