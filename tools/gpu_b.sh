#!/bin/bash
# round-2 GPU call B: batch-affine bucket reduction — parity, then timing against the plain path
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/r02_b_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_b_pytest.txt
tail -4 gpurun_out/r02_b_pytest.txt
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 5 --warmup 2 --sweep none --no-cpu > gpurun_out/r02_b_bench_$name.json 2> gpurun_out/r02_b_bench_$name.err
  rc=$?
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02_b_bench_$name.json'))
    print('$name', 'ms', round(d['ms_per_step'],3), 'op', {k:round(v,3) for k,v in d['op_ms'].items()}, 'acc', round(d['roofline']['isolated']['k_accumulate_ms'],3), 'aff', round(d['roofline']['isolated']['k_batch_affine_ms'],3), 'ver', d['verified']['msm'], d['verified']['msm_e2e'], 'e2e', round(d['e2e']['ms_per_step'],2))
except Exception as e:
    print('$name failed rc=$rc', e); print(open('gpurun_out/r02_b_bench_$name.err').read()[-800:])
PY
}
run lv0 H2B_AFF_LEVELS=0
run lv1 H2B_AFF_LEVELS=1
run lv2 H2B_AFF_LEVELS=2
run lv3 H2B_AFF_LEVELS=3
run lv3_k16 H2B_AFF_LEVELS=3 H2B_BA_K=16
run lv3_k64 H2B_AFF_LEVELS=3 H2B_BA_K=64
run lv3_k48 H2B_AFF_LEVELS=3 H2B_BA_K=48
run lv2_k64 H2B_AFF_LEVELS=2 H2B_BA_K=64
run lv3_c3 H2B_AFF_LEVELS=3 H2B_BA_CTAS=3
