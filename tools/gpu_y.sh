#!/bin/bash
# round-2 GPU call Y: column weights carry 2^ml * hi + 1 (no doublings / T sum left in k_weighted_final): tests + per-kernel times
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -k "msm or commit or kzg or grouped or eip" > gpurun_out/r02_y_pytest.txt 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r02_y_pytest.txt
for k in 10 16 19; do
  echo "##### k=$k" >> gpurun_out/r02_y_ops.txt
  timeout 300 python tools/prof_ops.py $k >> gpurun_out/r02_y_ops.txt 2>&1
done
grep -E "#####|== .*MSM|k_rowcol|k_weighted" gpurun_out/r02_y_ops.txt
