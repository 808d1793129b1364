#!/bin/bash
# round-2 GPU call AE: the clean-built library once more through every GPU test and the smoke entry
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
ls -la halo2-lib_b200/libh2b200.so
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02_ae_pytest_all.txt 2>&1
echo "all pytest rc=$?"; tail -3 gpurun_out/r02_ae_pytest_all.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02_ae_smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02_ae_smoke.txt
