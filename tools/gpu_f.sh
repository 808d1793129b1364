#!/bin/bash
# round-2 GPU call F: window size for small domains (k = 10 .. 15) + the GPU tests added so far
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
out=gpurun_out/r02_f_window_sweep_small.txt
: > $out
for k in 10 12 13 14 15; do
  for d in -3 -2 -1 0 1 2; do
    c=$((k + d))
    echo "##### k=$k c=$c" >> $out
    H2B_MSM_C=$c timeout 300 python tools/prof_ops.py $k 2>&1 | grep -E "== MSM" >> $out
  done
done
cat $out | paste - - - | awk '{print $2,$3,$7,$8,$15,$16}'
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "affine or eip196" 2>&1 | tail -5
