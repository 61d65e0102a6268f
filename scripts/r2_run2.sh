set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2_run2.log; : > $O
python scripts/quick_decode.py 4b 4096 64 2>&1 | tail -1 >> $O
for inf in 1 2 3; do QWEN_MEGA_INFLIGHT=$inf python scripts/quick_decode.py 4b 4096 64 2>&1 | tail -1 | sed "s/^/inflight $inf: /" >> $O; done
for ns in 6 7; do for inf in 0 2 3; do QWEN_MEGA_NSLOT=$ns QWEN_MEGA_INFLIGHT=$inf python scripts/quick_decode.py 4b 4096 64 2>&1 | tail -1 | sed "s/^/nslot $ns inflight $inf: /" >> $O; done; done
python scripts/phase_profile.py 4b 4096 2>&1 | grep -v "^\[" >> $O
QWEN_MEGA_INFLIGHT=2 python scripts/phase_profile.py 4b 4096 2>&1 | grep -v "^\[" >> $O
python scripts/quick_decode.py 8b 4096 32 2>&1 | tail -1 >> $O
python -m pytest tests/test_gpu_parity.py -x -q -k "golden_micro or real_layer or matmul or deterministic or greedy_256 or logits_and_kv" 2>&1 | tail -5 >> $O
