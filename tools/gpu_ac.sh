#!/bin/bash
# round-2 GPU call AC: final 1-GPU regression: every GPU test, smoke, the default bench line (all sweep configs + CPU baseline)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
if [ "$1" != "bench-only" ]; then
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r02_ac_pytest_all.txt 2>&1
echo "all pytest rc=$?"; tail -4 gpurun_out/r02_ac_pytest_all.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02_ac_smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02_ac_smoke.txt
fi
T0=$(date +%s); timeout 1500 python bench.py > gpurun_out/r02_ac_bench_default.json 2> gpurun_out/r02_ac_bench_default.err; echo "bench rc=$? wall $(( $(date +%s) - T0 )) s"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_ac_bench_default.json'))
print('ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],2), 'hostbuf', round(d['e2e_host_buffers']['ms_per_step'],2), 'frac', round(d['roofline']['frac'],5), 'iso', round(d['roofline']['isolated']['frac'],5), 'launches', d['gpu_launches'])
print('cpu', round(d['cpu_baseline']['value']), round(d['cpu_baseline']['step_ms'],1), d['cpu_baseline']['cores'])
for k,v in d['extra']['configs'].items(): print(k, v.get('k'), round(v.get('create_proof_schedule_ms',0),3), (v.get('e2e_resident_proof') or {}).get('ms_per_proof'), v.get('verified',{}).get('ok'), v.get('error'))
print('clocks', d['clocks'], 'verified', d['verified']['all_ranks_ok'])
PY
[ "$1" = "bench-only" ] || timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_ac_bench_reference.json 2> gpurun_out/r02_ac_bench_reference.err; echo "reference arm rc=$?"; cut -c1-300 gpurun_out/r02_ac_bench_reference.json
