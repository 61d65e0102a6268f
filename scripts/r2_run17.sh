cd $GRAFT_REPO_ROOT
O=gpurun_out/r2_run17.log; : > $O
for late in 0 1; do for pub in 0 1; do
QWEN_PW_LATE=$late QWEN_PW_PUB=$pub timeout 200 python scripts/quick_decode.py 4b 4096 64 2>&1 | tail -1 | sed "s/^/late $late pub $pub: /" >> $O
done; done
QWEN_PW_LATE=1 QWEN_MEGA_L2AHEAD=3 timeout 200 python scripts/quick_decode.py 4b 4096 64 2>&1 | tail -1 | sed "s/^/late 1 l2ahead 3: /" >> $O
QWEN_PW_LATE=1 QWEN_PW_PUB=1 timeout 200 python scripts/phase_profile.py 4b 4096 2>&1 | grep -v "^\[" >> $O
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "golden_micro or logits_and_kv or real_layer or deterministic" 2>&1 | tail -3 >> $O
