cd $GRAFT_REPO_ROOT
O=gpurun_out/r2_run35.log; : > $O
run() { timeout 200 python scripts/quick_decode.py 4b 4096 64 2>&1 | tail -1 | sed "s/^/$1: /" >> $O; }
QWEN_MEGA_L2MODE=0 run "17 warps, 96 regs, bulk prefetch only"
QWEN_MEGA_L2MODE=3 QWEN_MEGA_L2WIN_KB=256 run "warp-16 touches only, win 256"
QWEN_MEGA_L2MODE=3 QWEN_MEGA_L2WIN_KB=512 run "warp-16 touches only, win 512"
QWEN_MEGA_L2MODE=4 QWEN_MEGA_L2WIN_KB=128 run "bulk prefetch + warp-16 touches, win 128"
QWEN_MEGA_L2MODE=4 QWEN_MEGA_L2WIN_KB=256 run "bulk prefetch + warp-16 touches, win 256"
QWEN_MEGA_L2MODE=4 QWEN_MEGA_L2WIN_KB=384 run "bulk prefetch + warp-16 touches, win 384"
QWEN_MEGA_L2MODE=4 QWEN_MEGA_L2WIN_KB=256 QWEN_MEGA_L2BURST=8 run "bulk prefetch + warp-16 touches, win 256 burst 8"
QWEN_MEGA_L2MODE=4 QWEN_MEGA_L2WIN_KB=256 timeout 200 python scripts/phase_profile.py 4b 4096 2>&1 | grep -v "^\[" | grep -v skew >> $O
