set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2_run11.log; : > $O
for a in 2 0 1 3 4; do QWEN_MEGA_L2AHEAD=$a timeout 200 python scripts/quick_decode.py 4b 4096 64 2>&1 | tail -1 | sed "s/^/l2ahead $a: /" >> $O; done
timeout 200 python scripts/phase_profile.py 4b 4096 2>&1 | grep -v "^\[" >> $O
QWEN_MEGA_MODE=1 timeout 200 python scripts/phase_profile.py 4b 4096 2>&1 | grep -v "^\[" | grep -v skew >> $O
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "golden_micro or logits_and_kv or real_layer or deterministic or greedy_256" 2>&1 | tail -4 >> $O
