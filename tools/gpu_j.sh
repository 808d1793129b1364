#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_j_pytest_all.txt 2>&1
echo "all pytest rc=$?"; tail -30 gpurun_out/r02_j_pytest_all.txt
