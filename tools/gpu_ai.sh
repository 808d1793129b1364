#!/bin/bash
# round-2 GPU call AI: the resident prover against the committed golden proof (tests/golden/prover_k5.json)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_prover.py -m gpu -q -k "golden or oracle_prover or commitments_and_quotient" > gpurun_out/r02_ai_pytest.txt 2>&1
echo "pytest rc=$?"; tail -12 gpurun_out/r02_ai_pytest.txt
