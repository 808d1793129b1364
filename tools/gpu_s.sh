#!/bin/bash
# round-2 GPU call S (N GPUs): transforms in the background of latency-bound MSM phases (ntt.max_ctas_per_sm) vs in one block
N=${1:-4}
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
run() {
  name=$1; shift
  env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 8 --warmup 3 --sweep none --no-cpu > gpurun_out/r02_s_n${N}_$name.json 2> gpurun_out/r02_s_n${N}_$name.err
  rc=$?
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02_s_n${N}_$name.json'))
    print('N=$N $name', 'ms', round(d['ms_per_step'],3), 'ovl', round(d.get('create_proof_schedule_ms_ntt_overlap',0),3), 'seq', round(d.get('create_proof_schedule_ms_no_ntt_overlap',0),3), 'ver', d['verified']['msm'], d['verified']['all_ranks_ok'], 'e2e', round(d['e2e']['ms_per_step'],2))
except Exception as e:
    print('N=$N $name failed rc=$rc', e); print(open('gpurun_out/r02_s_n${N}_$name.err').read()[-1500:])
PY
}
run bg1 H2B_BENCH_BG_NTT=1
run bg2 H2B_BENCH_BG_NTT=2
