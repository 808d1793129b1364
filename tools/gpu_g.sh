#!/bin/bash
# round-2 GPU call G: resident prover path — tests, then the full bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_prover.py -m gpu -x -q > gpurun_out/r02_g_pytest_prover.txt 2>&1
echo "prover pytest rc=$?"; tail -25 gpurun_out/r02_g_pytest_prover.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_g_bench.json 2> gpurun_out/r02_g_bench.err
echo "bench rc=$?"; tail -5 gpurun_out/r02_g_bench.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r02_g_bench.json'))
    print('ms', round(d['ms_per_step'],3), 'e2e', {k:v for k,v in d['e2e'].items() if k!='path'})
    print('e2e_host', {k:v for k,v in d['e2e_host_buffers'].items() if k!='path'})
    print('verified', d['verified'])
    for k,v in d['extra']['configs'].items(): print(k, v.get('create_proof_schedule_ms'), v.get('msm_only_ms'), v.get('verified',{}).get('ok'), v.get('error'))
except Exception as e: print('bench parse failed', e)
PY
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_g_pytest_all.txt 2>&1
echo "all pytest rc=$?"; tail -5 gpurun_out/r02_g_pytest_all.txt
