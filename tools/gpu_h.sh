#!/bin/bash
# round-2 GPU call H: resident prover with side-queue transforms and batched evaluations
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_prover.py tests/test_gpu_quotient.py -m gpu -x -q > gpurun_out/r02_h_pytest.txt 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/r02_h_pytest.txt
H2B_PROVER_TRACE=1 timeout 900 python bench.py --steps 3 --warmup 2 --sweep none --no-cpu > gpurun_out/r02_h_bench_trace.json 2> gpurun_out/r02_h_bench_trace.err
echo "trace rc=$?"; grep "prover trace" gpurun_out/r02_h_bench_trace.err | tail -2
timeout 900 python bench.py --steps 10 --warmup 3 --sweep none --no-cpu > gpurun_out/r02_h_bench.json 2> gpurun_out/r02_h_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r02_h_bench.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r02_h_bench.json'))
    print('ms', round(d['ms_per_step'],3), 'e2e', {k:v for k,v in d['e2e'].items() if k!='path'})
    print('verified', {k:v for k,v in d['verified'].items() if k not in ('method','ntt')})
except Exception as e: print('bench parse failed', e)
PY
