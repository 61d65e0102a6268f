cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -q 2>&1 | grep -v "^\[" | grep -B5 -A40 "def test_staged\|Error" | head -120 > gpurun_out/r2_parity_full2.log
timeout 1500 python -m pytest tests/test_gpu_strict.py -x -q -s 2>&1 | grep -v "^\[Params\]\|^\[Weights\]\|^\[Forward" | tail -40 > gpurun_out/r2_strict.log
