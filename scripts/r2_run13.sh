cd $GRAFT_REPO_ROOT
O=gpurun_out/r2_run13.log; : > $O
timeout 200 python scripts/quick_decode.py 4b 4096 64 2>&1 | tail -1 >> $O
QWEN_MEGA_L2AHEAD=0 timeout 200 python scripts/quick_decode.py 4b 4096 64 2>&1 | tail -1 | sed "s/^/l2ahead 0: /" >> $O
QWEN3_LIB_PATH=qwen3.c_b200/lib_prof/libqwen3.so timeout 200 python scripts/unit_profile.py 4b 4096 2>&1 | grep -v "^\[" >> $O
timeout 200 python scripts/phase_profile.py 4b 4096 2>&1 | grep -v "^\[" >> $O
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "golden_micro or logits_and_kv or real_layer or deterministic or greedy_256" 2>&1 | tail -4 >> $O
