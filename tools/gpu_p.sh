#!/bin/bash
# round-2 GPU call P: high-priority tail streams + split row/column weights: tests, per-kernel times, timeline, bench variants
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_prover.py -m gpu -q > gpurun_out/r02_p_pytest.txt 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r02_p_pytest.txt
for k in 16 19; do
  for tp in 1 0; do
    echo "##### k=$k H2B_TAIL_PRIORITY=$tp" >> gpurun_out/r02_p_ops.txt
    H2B_TAIL_PRIORITY=$tp timeout 300 python tools/prof_ops.py $k >> gpurun_out/r02_p_ops.txt 2>&1
  done
done
grep -E "#####|== .*MSM" gpurun_out/r02_p_ops.txt
timeout 300 python tools/timeline.py --out gpurun_out/r02_p_timeline.csv > gpurun_out/r02_p_timeline.txt 2>&1; head -30 gpurun_out/r02_p_timeline.txt
run() {
  name=$1; sweep=$2; shift; shift
  env "$@" timeout 600 python bench.py --steps 5 --warmup 3 --sweep $sweep --no-cpu > gpurun_out/r02_p_bench_$name.json 2> gpurun_out/r02_p_bench_$name.err
  rc=$?
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02_p_bench_$name.json'))
    print('$name', 'ms', round(d['ms_per_step'],3), 'seq', round(d.get('create_proof_schedule_ms_no_ntt_overlap',0),3), 'ver', d['verified']['msm'], d['verified']['msm_e2e'], 'e2e', round(d['e2e']['ms_per_step'],2), 'hostbuf', round(d['e2e_host_buffers']['ms_per_step'],2), 'launches', d['gpu_launches'])
    for k,v in d.get('extra',{}).get('configs',{}).items(): print('   ', k, 'k', v.get('k'), 'ms', round(v.get('create_proof_schedule_ms',0),3), 'ok', v.get('verified',{}).get('ok'), v.get('error'))
except Exception as e:
    print('$name failed rc=$rc', e); print(open('gpurun_out/r02_p_bench_$name.err').read()[-800:])
PY
}
run tp1_g1 1,2 H2B_TAIL_PRIORITY=1 H2B_MSM_GROUP=1
run tp1_gdef 1,2,4 H2B_TAIL_PRIORITY=1
run tp0_g1 none H2B_TAIL_PRIORITY=0 H2B_MSM_GROUP=1
run tp1_g2 none H2B_TAIL_PRIORITY=1 H2B_MSM_GROUP=2
