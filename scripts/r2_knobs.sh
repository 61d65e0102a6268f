cd $GRAFT_REPO_ROOT
O=gpurun_out/r2_knobs.log; : > $O
run() { timeout 200 python scripts/quick_decode.py 4b 4096 128 2>&1 | tail -1 | sed "s/^/$1: /" >> $O; }
run "default"
QWEN_MEGA_STAGE=0 run "STAGE=0 (per-warp stores + fence)"
QWEN_MEGA_STAGE_MASK=12 run "STAGE_MASK=12 (w13,w2 staged; qkv,wo per-warp)"
QWEN_MEGA_STAGE_MASK=19 run "STAGE_MASK=19 (qkv,wo,cls staged)"
QWEN_MEGA_STAGE_MASK=27 run "STAGE_MASK=27 (all but w13)"
QWEN_PW_LATE=1 run "PW_LATE=1"
QWEN_PW_PUB=1 run "PW_PUB=1"
QWEN_MEGA_L2AHEAD=1 run "L2AHEAD=1"
QWEN_MEGA_L2AHEAD=3 run "L2AHEAD=3"
run "default again"
cat $O
