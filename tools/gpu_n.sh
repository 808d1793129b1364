#!/bin/bash
# round-2 GPU call N: hand-written radix sort of the lookup permutation (tests, sanitizer, timings) + the tests that failed in call M
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_quotient.py tests/test_gpu_prover.py tests/test_cpp_mirror.py tests/test_gpu_parity.py -m gpu -q > gpurun_out/r02_n_pytest.txt 2>&1
echo "pytest rc=$?"; tail -12 gpurun_out/r02_n_pytest.txt
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_quotient.py -m gpu -q -k "permute_expression_pair and not 18 and not 17" > gpurun_out/r02_n_racecheck_sort.txt 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|ERROR SUMMARY" gpurun_out/r02_n_racecheck_sort.txt | tail -4
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_quotient.py -m gpu -q -k "permute_expression_pair and not 18 and not 17" > gpurun_out/r02_n_memcheck_sort.txt 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r02_n_memcheck_sort.txt | tail -4
timeout 900 python tools/prof_quotient.py 19 > gpurun_out/r02_n_next_rows.txt 2>&1
echo "prof rc=$?"; grep -i "permute\|g_to_lagrange\|quotient_graph" gpurun_out/r02_n_next_rows.txt | cut -c1-220
