#!/bin/bash
# round-2 GPU call AH: resident prover against the ORACLE prover (byte for byte), C++ twin, and the bench line after the top-coefficient fix
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_prover.py -m gpu -q > gpurun_out/r02_ah_pytest_prover.txt 2>&1
echo "pytest rc=$?"; tail -25 gpurun_out/r02_ah_pytest_prover.txt
timeout 600 python bench.py --steps 10 --warmup 3 --sweep 2 --no-cpu > gpurun_out/r02_ah_bench.json 2> gpurun_out/r02_ah_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_ah_bench.json'))
print('ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), d['verified']['msm_e2e'], d['verified']['e2e_quotient_identity'], d['verified']['all_ranks_ok'])
for k,v in d['extra']['configs'].items(): print(k, v.get('k'), round(v.get('create_proof_schedule_ms',0),3), (v.get('e2e_resident_proof') or {}).get('ms_per_proof'), v.get('verified',{}).get('ok'), v.get('error'))
PY
