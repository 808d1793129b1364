#!/bin/bash
# round-2 GPU call L: full GPU suite (everything new since call G) + batch-affine with per-thread safegcd inversion
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02_l_pytest_all.txt 2>&1
echo "all pytest rc=$?"; tail -25 gpurun_out/r02_l_pytest_all.txt
run() {
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 5 --warmup 2 --sweep none --no-cpu > gpurun_out/r02_l_bench_$name.json 2> gpurun_out/r02_l_bench_$name.err
  rc=$?
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02_l_bench_$name.json'))
    print('$name', 'ms', round(d['ms_per_step'],3), 'msm_u', round(d['op_ms']['msm_uniform'],3), 'msm_w', round(d['op_ms']['msm_witness'],3), 'acc', round(d['roofline']['isolated']['k_accumulate_ms'],3), 'aff', round(d['roofline']['isolated']['k_batch_affine_ms'],3), 'ver', d['verified']['msm'], d['verified']['msm_e2e'], 'e2e', round(d['e2e']['ms_per_step'],2))
except Exception as e:
    print('$name failed rc=$rc', e); print(open('gpurun_out/r02_l_bench_$name.err').read()[-800:])
PY
}
run plain H2B_AFF_LEVELS=0
run pt3_k32 H2B_AFF_LEVELS=3 H2B_BA_PT=1 H2B_BA_K=32
run pt3_k64 H2B_AFF_LEVELS=3 H2B_BA_PT=1 H2B_BA_K=64
run pt2_k64 H2B_AFF_LEVELS=2 H2B_BA_PT=1 H2B_BA_K=64
run pt1_k64 H2B_AFF_LEVELS=1 H2B_BA_PT=1 H2B_BA_K=64
run pt3_k64_c5 H2B_AFF_LEVELS=3 H2B_BA_PT=1 H2B_BA_K=64 H2B_BA_CTAS=5
