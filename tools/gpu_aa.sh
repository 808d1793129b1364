#!/bin/bash
# round-2 GPU call AA: compiled vs Python host side of the resident prover, timed; bench line after the affine-form change
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python tools/bench_cpp_prover.py 19 1 0 10 > gpurun_out/r02_aa_cpp_prover_k19.txt 2>&1; echo "cpp k19 rc=$?"; tail -3 gpurun_out/r02_aa_cpp_prover_k19.txt
timeout 900 python tools/bench_cpp_prover.py 16 8 2 10 > gpurun_out/r02_aa_cpp_prover_k16.txt 2>&1; echo "cpp k16 rc=$?"; tail -3 gpurun_out/r02_aa_cpp_prover_k16.txt
timeout 900 python tools/bench_cpp_prover.py 14 1 0 20 > gpurun_out/r02_aa_cpp_prover_k14.txt 2>&1; echo "cpp k14 rc=$?"; tail -3 gpurun_out/r02_aa_cpp_prover_k14.txt
timeout 900 python bench.py --steps 10 --warmup 3 --sweep none --no-cpu > gpurun_out/r02_aa_bench.json 2> gpurun_out/r02_aa_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_aa_bench.json'))
print('ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), d['verified']['msm_e2e'], d['verified']['e2e_quotient_identity'])
PY
