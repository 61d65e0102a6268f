cd $GRAFT_REPO_ROOT
O=gpurun_out/r2_run28.log; : > $O
for m in 31 0 20 4 21 28; do QWEN_MEGA_STAGE_MASK=$m timeout 200 python scripts/quick_decode.py 4b 4096 64 2>&1 | tail -1 | sed "s/^/stage mask $m: /" >> $O; done
timeout 200 python scripts/phase_profile.py 4b 4096 2>&1 | grep -v "^\[" | grep -v skew >> $O
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "golden_micro or logits_and_kv or real_layer or deterministic or long_context" 2>&1 | tail -3 >> $O
