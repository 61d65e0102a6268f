cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_strict.py -x -q -s -k "config2 or config1 or patched" 2>&1 | grep -v "^\[Params\]\|^\[Weights\]\|^\[Forward" | tail -30 > gpurun_out/r2_strict2.log
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "staged or golden_micro or real_layer" 2>&1 | tail -4 > gpurun_out/r2_staged.log
timeout 600 python bench.py --steps 64 --warmup 5 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err
