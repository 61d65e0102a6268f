set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2_run3.log; : > $O
QWEN3_LIB_PATH=qwen3.c_b200/lib_prof/libqwen3.so python scripts/unit_profile.py 4b 4096 2>&1 | grep -v "^\[" >> $O
QWEN_MEGA_NSLOT=7 QWEN3_LIB_PATH=qwen3.c_b200/lib_prof/libqwen3.so python scripts/unit_profile.py 4b 4096 2>&1 | grep -v "^\[" >> $O
python scripts/quick_decode.py 4b 4096 64 2>&1 | tail -1 >> $O
