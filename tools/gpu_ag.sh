#!/bin/bash
# round-2 GPU call AG: C++ mirror with grouped commitments and set_option
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cpp_mirror.py -m gpu -q > gpurun_out/r02_ag_pytest.txt 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/r02_ag_pytest.txt
