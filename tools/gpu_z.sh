#!/bin/bash
# round-2 GPU call Z: the compiled (C++) host side of the resident prover against the Python one, byte for byte
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_prover.py tests/test_cpp_mirror.py -m gpu -q -x > gpurun_out/r02_z_pytest.txt 2>&1
echo "pytest rc=$?"; tail -30 gpurun_out/r02_z_pytest.txt
