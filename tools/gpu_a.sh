#!/bin/bash
# round-2 GPU call A: sanity of the restructured bench + the GPU test suite
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r02_a_smi.txt 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_a_bench.json 2> gpurun_out/r02_a_bench.err
echo "bench rc=$?" >> gpurun_out/r02_a_bench.err
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_a_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_a_pytest.txt
tail -5 gpurun_out/r02_a_pytest.txt
tail -3 gpurun_out/r02_a_bench.err
head -c 1500 gpurun_out/r02_a_bench.json
