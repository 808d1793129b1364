#!/bin/bash
# round-2 GPU call U: smoke, randomised soak and sanitizer over the round-2 kernels, then the default bench line (with CPU baseline)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02_u_smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02_u_smoke.txt
timeout 400 python tools/soak.py 150 > gpurun_out/r02_u_soak.txt 2>&1; echo "soak rc=$?"; tail -3 gpurun_out/r02_u_soak.txt
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -k "group_pipeline or edge_cases or hot_bucket" > gpurun_out/r02_u_memcheck_msm.txt 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r02_u_memcheck_msm.txt | tail -3
timeout 1200 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -q -k "group_pipeline and (0 or 16)" > gpurun_out/r02_u_racecheck_msm.txt 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/r02_u_racecheck_msm.txt | tail -3
timeout 1200 python bench.py > gpurun_out/r02_u_bench_default.json 2> gpurun_out/r02_u_bench_default.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_u_bench_default.json'))
print('ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],2), 'hostbuf', round(d['e2e_host_buffers']['ms_per_step'],2), 'frac', d['roofline']['frac'], 'iso', d['roofline']['isolated']['frac'], 'cols/launch', d['roofline']['msm_columns_per_launch'])
print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['step_ms'], d['cpu_baseline']['cores'], d['cpu_baseline']['op_ms'])
for k,v in d['extra']['configs'].items(): print(k, v.get('k'), round(v.get('create_proof_schedule_ms',0),3), v.get('verified',{}).get('ok'), v.get('error'))
print('clocks', d['clocks'])
PY
