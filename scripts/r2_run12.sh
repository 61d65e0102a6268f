cd $GRAFT_REPO_ROOT
O=gpurun_out/r2_run12.log; : > $O
QWEN3_LIB_PATH=qwen3.c_b200/lib_prof/libqwen3.so timeout 200 python scripts/unit_profile.py 4b 4096 2>&1 | grep -v "^\[" >> $O
