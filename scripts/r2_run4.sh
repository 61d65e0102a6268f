set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2_run5.log; : > $O
python scripts/quick_decode.py 4b 4096 64 2>&1 | tail -1 >> $O
QWEN_MEGA_NSLOT=7 python scripts/quick_decode.py 4b 4096 64 2>&1 | tail -1 >> $O
QWEN3_LIB_PATH=qwen3.c_b200/lib_prof/libqwen3.so python scripts/unit_profile.py 4b 4096 2>&1 | grep -v "^\[" >> $O
python scripts/phase_profile.py 4b 4096 2>&1 | grep -v "^\[" >> $O
python -m pytest tests/test_gpu_parity.py -x -q -k "golden_micro or real_layer or deterministic or greedy_256 or logits_and_kv or staged" 2>&1 | tail -5 >> $O
