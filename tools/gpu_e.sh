#!/bin/bash
# round-2 GPU call E: window size vs shard size (what an N-GPU shard of the k=19 column sees) + kernel breakdown
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
out=gpurun_out/r02_e_window_sweep.txt
: > $out
for k in 16 17 18 19; do
  for c in 11 12 13 14 15 16 17; do
    echo "##### k=$k c=$c" >> $out
    H2B_MSM_C=$c timeout 300 python tools/prof_ops.py $k 2>&1 | grep -v "^$" | head -24 >> $out
  done
done
grep -E "#####|== MSM" $out
