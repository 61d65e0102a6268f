cd $GRAFT_REPO_ROOT
O=gpurun_out/r2_run31.log; : > $O
QWEN_MEGA_L2MODE=0 timeout 200 python scripts/quick_decode.py 4b 4096 64 2>&1 | tail -1 | sed "s/^/bulk prefetch (old): /" >> $O
for W in 128 256 384 512 640; do
QWEN_MEGA_L2WIN_KB=$W timeout 200 python scripts/quick_decode.py 4b 4096 64 2>&1 | tail -1 | sed "s/^/near win $W KB: /" >> $O
done
QWEN_MEGA_L2GROUPS=8 timeout 200 python scripts/quick_decode.py 4b 4096 64 2>&1 | tail -1 | sed "s/^/near groups 8: /" >> $O
QWEN_MEGA_L2GROUPS=4 timeout 200 python scripts/quick_decode.py 4b 4096 64 2>&1 | tail -1 | sed "s/^/near groups 4: /" >> $O
timeout 200 python scripts/phase_profile.py 4b 4096 2>&1 | grep -v "^\[" | grep -v skew >> $O
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "golden_micro or logits_and_kv or real_layer or deterministic or greedy_256 or staged" 2>&1 | tail -3 >> $O
timeout 200 python scripts/quick_decode.py 1.7b 512 64 2>&1 | tail -1 >> $O
timeout 200 python scripts/quick_decode.py 0.6b 128 64 2>&1 | tail -1 >> $O
