set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
python scripts/quick_decode.py 4b 4096 64 > gpurun_out/r2_q0.log 2>&1
for ns in 5 7; do for mode in 0 1; do
  echo "== nslot $ns mode $mode" >> gpurun_out/r2_phase0.log
  QWEN_MEGA_NSLOT=$ns QWEN_MEGA_MODE=$mode timeout 300 python scripts/phase_profile.py 4b 4096 >> gpurun_out/r2_phase0.log 2>&1
done; done
timeout 120 scripts/ubench/attn > gpurun_out/r2_ubench_attn.log 2>&1
timeout 200 scripts/ubench/ubench > gpurun_out/r2_ubench.log 2>&1
