#!/bin/bash
# round-2 GPU call V: resident prover for every BASELINE shape (multi-column, lookup-advice, lookup-less): tests + bench sweep
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_prover.py -m gpu -q -x > gpurun_out/r02_v_pytest_prover.txt 2>&1
echo "pytest rc=$?"; tail -25 gpurun_out/r02_v_pytest_prover.txt
timeout 1500 python bench.py --steps 5 --warmup 3 --sweep 1,2,4 --no-cpu > gpurun_out/r02_v_bench.json 2> gpurun_out/r02_v_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/r02_v_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_v_bench.json'))
print('ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],2), d['verified']['msm_e2e'], d['verified']['e2e_quotient_identity'])
for k,v in d['extra']['configs'].items(): print(k, v.get('k'), round(v.get('create_proof_schedule_ms',0),3), v.get('e2e_resident_proof'), v.get('verified',{}).get('ok'), v.get('error'))
PY
