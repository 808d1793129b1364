#!/bin/bash
# round-2 GPU call M: full GPU suite + grouped MSM pipelines (msm.batch_group) at k=19 and on the small sweep configs
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r02_m_pytest_all.txt 2>&1
echo "all pytest rc=$?"; tail -15 gpurun_out/r02_m_pytest_all.txt
run() {
  name=$1; sweep=$2; shift; shift
  env "$@" timeout 600 python bench.py --steps 5 --warmup 3 --sweep $sweep --no-cpu > gpurun_out/r02_m_bench_$name.json 2> gpurun_out/r02_m_bench_$name.err
  rc=$?
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02_m_bench_$name.json'))
    print('$name', 'ms', round(d['ms_per_step'],3), 'seq', round(d.get('create_proof_schedule_ms_no_ntt_overlap',0),3), 'msm_u', round(d['op_ms']['msm_uniform'],3), 'msm_w', round(d['op_ms']['msm_witness'],3), 'ver', d['verified']['msm'], d['verified']['msm_e2e'], 'e2e', round(d['e2e']['ms_per_step'],2), 'hostbuf', round(d['e2e_host_buffers']['ms_per_step'],2), 'launches', d['gpu_launches'])
    for k,v in d.get('extra',{}).get('configs',{}).items(): print('   ', k, 'k', v.get('k'), 'ms', round(v.get('create_proof_schedule_ms',0),3), 'ok', v.get('verified',{}).get('ok'), v.get('error'))
except Exception as e:
    print('$name failed rc=$rc', e); print(open('gpurun_out/r02_m_bench_$name.err').read()[-800:])
PY
}
run g1 1,2 H2B_MSM_GROUP=1
run gdef 1,2,4,5 H2B_MSM_GROUP=0
run g2 none H2B_MSM_GROUP=2
run g4 1,2 H2B_MSM_GROUP=4
run g16 1,2 H2B_MSM_GROUP=16
