set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2_run9.log; : > $O
for mode in 1 2 3; do
  echo "== pw mode $mode (bit0: no dot products, bit1: no weight traffic)" >> $O
  QWEN_MEGA_MODE=$mode timeout 200 python scripts/phase_profile.py 4b 4096 2>&1 | grep -v "^\[" >> $O
done
