set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2_run6.log; : > $O
python scripts/quick_decode.py small 128 8 2>&1 | tail -2 >> $O
QWEN3_LIB_PATH=qwen3.c_b200/lib_b/libqwen3.so python scripts/quick_decode.py small 128 8 2>&1 | tail -2 >> $O
timeout 300 compute-sanitizer --tool memcheck python scripts/quick_decode.py small 128 2 2>&1 | grep -v "^\[" | head -40 >> $O
