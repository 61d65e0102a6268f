cd $GRAFT_REPO_ROOT
O=gpurun_out/r2_run27.log; : > $O
for st in 1 0; do QWEN_MEGA_STAGE=$st timeout 200 python scripts/quick_decode.py 4b 4096 64 2>&1 | tail -1 | sed "s/^/stage $st: /" >> $O; done
QWEN_MEGA_STAGE=0 timeout 200 python scripts/phase_profile.py 4b 4096 2>&1 | grep -v "^\[" >> $O
QWEN_MEGA_STAGE=0 timeout 200 python scripts/quick_decode.py 8b 4096 32 2>&1 | tail -1 | sed "s/^/stage 0: /" >> $O
