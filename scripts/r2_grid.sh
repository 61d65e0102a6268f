cd $GRAFT_REPO_ROOT
O=gpurun_out/r2_grid.log; : > $O
for sh in "1.7b 512" "0.6b 128" "4b 4096"; do
  for g in 148 128 111 96 74; do
    QWEN_MEGA_GRID=$g timeout 200 python scripts/quick_decode.py $sh 64 2>&1 | tail -1 | sed "s/^/grid $g: /" >> $O
  done
done
cat $O
