#!/bin/bash
# round-2 GPU call C: where does the batch-affine reduction pay?  k = 21 and k = 23 against the plain path
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
run() {  # name, bench args..., env via H2B_*
  name=$1; shift
  timeout 900 python bench.py --steps 3 --warmup 2 --sweep none --no-cpu "$@" > gpurun_out/r02_c_bench_$name.json 2> gpurun_out/r02_c_bench_$name.err
  rc=$?
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02_c_bench_$name.json'))
    print('$name', 'ms', round(d['ms_per_step'],3), 'op', {k:round(v,3) for k,v in d['op_ms'].items()}, 'acc', round(d['roofline']['isolated']['k_accumulate_ms'],3), 'aff', round(d['roofline']['isolated']['k_batch_affine_ms'],3), 'ver', d['verified']['msm'], d['verified']['msm_e2e'], 'e2e', round(d['e2e']['ms_per_step'],2))
except Exception as e:
    print('$name failed rc=$rc', e); print(open('gpurun_out/r02_c_bench_$name.err').read()[-800:])
PY
}
H2B_AFF_LEVELS=0 run k21_lv0 --config 3 --k 21
H2B_AFF_LEVELS=3 run k21_lv3 --config 3 --k 21
H2B_AFF_LEVELS=3 H2B_BA_K=64 run k21_lv3_k64 --config 3 --k 21
H2B_AFF_LEVELS=0 run k23_lv0 --config 5
H2B_AFF_LEVELS=3 run k23_lv3 --config 5
H2B_AFF_LEVELS=3 H2B_BA_K=64 run k23_lv3_k64 --config 5
H2B_AFF_LEVELS=3 H2B_BA_K=128 run k23_lv3_k128 --config 5
H2B_AFF_LEVELS=2 H2B_BA_K=64 run k23_lv2_k64 --config 5
