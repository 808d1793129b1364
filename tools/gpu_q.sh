#!/bin/bash
# round-2 GPU call Q: three stream priority levels (tails > lanes / main > transforms)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
run() {
  name=$1; sweep=$2; shift; shift
  env "$@" timeout 600 python bench.py --steps 5 --warmup 3 --sweep $sweep --no-cpu > gpurun_out/r02_q_bench_$name.json 2> gpurun_out/r02_q_bench_$name.err
  rc=$?
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02_q_bench_$name.json'))
    print('$name', 'ms', round(d['ms_per_step'],3), 'seq', round(d.get('create_proof_schedule_ms_no_ntt_overlap',0),3), 'ver', d['verified']['msm'], d['verified']['msm_e2e'], 'e2e', round(d['e2e']['ms_per_step'],2), 'hostbuf', round(d['e2e_host_buffers']['ms_per_step'],2), 'launches', d['gpu_launches'])
    for k,v in d.get('extra',{}).get('configs',{}).items(): print('   ', k, 'k', v.get('k'), 'ms', round(v.get('create_proof_schedule_ms',0),3), 'ok', v.get('verified',{}).get('ok'), v.get('error'))
except Exception as e:
    print('$name failed rc=$rc', e); print(open('gpurun_out/r02_q_bench_$name.err').read()[-800:])
PY
}
run prio_gdef 1,2,4 H2B_LANE_PRIORITY=1
run flat_gdef none H2B_LANE_PRIORITY=0 H2B_BENCH_PRIORITY=0
run prio_g1 none H2B_LANE_PRIORITY=1 H2B_MSM_GROUP=1
run prio_g4 none H2B_LANE_PRIORITY=1 H2B_MSM_GROUP=4
timeout 300 python tools/timeline.py --out gpurun_out/r02_q_timeline.csv > gpurun_out/r02_q_timeline.txt 2>&1; head -3 gpurun_out/r02_q_timeline.txt
