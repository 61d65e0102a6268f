cd $GRAFT_REPO_ROOT
O=gpurun_out/r2_run32.log; : > $O
run() { timeout 200 python scripts/quick_decode.py 4b 4096 64 2>&1 | tail -1 | sed "s/^/$1: /" >> $O; }
QWEN_MEGA_L2MODE=0 run "bulk prefetch only"
QWEN_MEGA_L2MODE=1 QWEN_MEGA_L2WIN_KB=192 run "both, win 192"
QWEN_MEGA_L2MODE=1 QWEN_MEGA_L2WIN_KB=256 run "both, win 256"
QWEN_MEGA_L2MODE=1 QWEN_MEGA_L2WIN_KB=384 run "both, win 384"
QWEN_MEGA_L2MODE=1 QWEN_MEGA_L2WIN_KB=256 QWEN_MEGA_L2BURST=8 run "both, win 256 burst 8"
QWEN_MEGA_L2MODE=1 QWEN_MEGA_L2WIN_KB=256 QWEN_MEGA_L2BURST=2 run "both, win 256 burst 2"
QWEN_MEGA_L2MODE=1 QWEN_MEGA_L2WIN_KB=256 QWEN_MEGA_L2BURST=16 QWEN_MEGA_L2GROUPS=16 run "both, win 256 burst 16"
QWEN_MEGA_L2MODE=2 QWEN_MEGA_L2WIN_KB=256 QWEN_MEGA_L2BURST=8 run "near only, win 256 burst 8"
QWEN_MEGA_L2MODE=2 QWEN_MEGA_L2WIN_KB=512 QWEN_MEGA_L2BURST=8 run "near only, win 512 burst 8"
QWEN_MEGA_L2MODE=2 QWEN_MEGA_L2WIN_KB=512 QWEN_MEGA_L2BURST=16 run "near only, win 512 burst 16"
QWEN_MEGA_L2MODE=1 QWEN_MEGA_L2WIN_KB=256 timeout 200 python scripts/phase_profile.py 4b 4096 2>&1 | grep -v "^\[" | grep -v skew >> $O
