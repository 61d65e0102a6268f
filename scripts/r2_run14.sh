cd $GRAFT_REPO_ROOT
O=gpurun_out/r2_run14.log; : > $O
timeout 120 python scripts/quick_decode.py small 128 8 2>&1 | tail -1 >> $O
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "golden_micro or logits_and_kv or real_layer or deterministic or greedy_256 or long_context or layered" 2>&1 | tail -6 >> $O
for a in 2 0 1 3; do QWEN_MEGA_L2AHEAD=$a timeout 200 python scripts/quick_decode.py 4b 4096 64 2>&1 | tail -1 | sed "s/^/l2ahead $a: /" >> $O; done
timeout 200 python scripts/phase_profile.py 4b 4096 2>&1 | grep -v "^\[" >> $O
QWEN3_LIB_PATH=qwen3.c_b200/lib_prof/libqwen3.so timeout 200 python scripts/unit_profile.py 4b 4096 2>&1 | grep -v "^\[" >> $O
timeout 200 python scripts/quick_decode.py 8b 4096 32 2>&1 | tail -1 >> $O
timeout 200 python scripts/quick_decode.py 1.7b 512 64 2>&1 | tail -1 >> $O
timeout 200 python scripts/quick_decode.py 0.6b 128 64 2>&1 | tail -1 >> $O
