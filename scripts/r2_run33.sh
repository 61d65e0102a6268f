cd $GRAFT_REPO_ROOT
O=gpurun_out/r2_run33.log; : > $O
run() { timeout 200 python scripts/quick_decode.py 4b 4096 64 2>&1 | tail -1 | sed "s/^/$1: /" >> $O; }
QWEN_MEGA_VERBOSE=1 QWEN_MEGA_L2MODE=1 timeout 200 python scripts/quick_decode.py 4b 4096 8 2>&1 | grep "near-L2" >> $O
QWEN_MEGA_L2MODE=0 run "bulk prefetch only"
QWEN_MEGA_L2MODE=1 QWEN_MEGA_L2WIN_KB=128 run "both, win 128"
QWEN_MEGA_L2MODE=1 QWEN_MEGA_L2WIN_KB=256 run "both, win 256"
QWEN_MEGA_L2MODE=1 QWEN_MEGA_L2WIN_KB=384 run "both, win 384"
QWEN_MEGA_L2MODE=1 QWEN_MEGA_L2WIN_KB=512 run "both, win 512"
QWEN_MEGA_L2MODE=1 QWEN_MEGA_L2WIN_KB=256 QWEN_MEGA_L2BURST=1 run "both, win 256 burst 1"
QWEN_MEGA_L2MODE=1 QWEN_MEGA_L2WIN_KB=256 QWEN_MEGA_L2BURST=4 run "both, win 256 burst 4"
QWEN_MEGA_L2MODE=2 QWEN_MEGA_L2WIN_KB=256 run "touches only, win 256"
QWEN_MEGA_L2MODE=2 QWEN_MEGA_L2WIN_KB=512 run "touches only, win 512"
QWEN_MEGA_L2MODE=2 QWEN_MEGA_L2WIN_KB=768 run "touches only, win 768"
QWEN_MEGA_L2MODE=1 QWEN_MEGA_L2WIN_KB=256 timeout 200 python scripts/phase_profile.py 4b 4096 2>&1 | grep -v "^\[" | grep -v skew >> $O
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "golden_micro or logits_and_kv or real_layer or deterministic or greedy_256 or staged" 2>&1 | tail -3 >> $O
