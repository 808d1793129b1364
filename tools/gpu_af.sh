#!/bin/bash
# round-2 GPU call AF: compute-sanitizer over the kernels added late in the round (GLV G1 FFT, folded column weights, prover kernels)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_srs.py -m gpu -q -k "not 16" > gpurun_out/r02_af_memcheck_srs.txt 2>&1
echo "memcheck srs rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r02_af_memcheck_srs.txt | tail -3
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_prover.py -m gpu -q -k "resident_proof and 8-1-0" > gpurun_out/r02_af_memcheck_prover.txt 2>&1
echo "memcheck prover rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r02_af_memcheck_prover.txt | tail -3
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -q -k "edge_cases or srs_vs_pippenger" > gpurun_out/r02_af_racecheck_msm.txt 2>&1
echo "racecheck msm rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/r02_af_racecheck_msm.txt | tail -3
