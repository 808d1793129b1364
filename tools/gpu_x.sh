#!/bin/bash
# round-2 GPU call X: G1 FFT with GLV scalar multiplications: SRS tests + timing
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_srs.py -m gpu -q > gpurun_out/r02_x_pytest_srs.txt 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r02_x_pytest_srs.txt
timeout 600 python tools/prof_quotient.py 19 > gpurun_out/r02_x_next_rows.txt 2>&1
echo "prof rc=$?"; grep -i "g_to_lagrange\|permute" gpurun_out/r02_x_next_rows.txt | cut -c1-200
