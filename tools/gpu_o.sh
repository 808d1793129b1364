#!/bin/bash
# round-2 GPU call O: NTT tile size by domain (H2B_NTT_TILE=11 = the former fixed 2^11 tile) and grouped MSM pipelines, per kernel
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -k "ntt or coset or kzg or transform" > gpurun_out/r02_o_pytest_ntt.txt 2>&1
echo "pytest ntt rc=$?"; tail -3 gpurun_out/r02_o_pytest_ntt.txt
for k in 14 16 19 20; do
  for tile in 0 11; do
    echo "##### k=$k H2B_NTT_TILE=$tile" >> gpurun_out/r02_o_ops.txt
    H2B_NTT_TILE=$tile timeout 300 python tools/prof_ops.py $k >> gpurun_out/r02_o_ops.txt 2>&1
  done
done
grep -E "#####|== " gpurun_out/r02_o_ops.txt
