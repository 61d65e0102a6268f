cd $GRAFT_REPO_ROOT
O=gpurun_out/r2_prefill_a3.log; : > $O
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "prefill_attention_kernels" 2>&1 | grep -v "^\[" | tail -30 >> $O
cat $O
