cd $GRAFT_REPO_ROOT
O=gpurun_out/r2_prefill_a2.log; : > $O
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "prefill_matches_reference_token_by_token" 2>&1 | grep -v "^\[" | tail -40 >> $O
cat $O
