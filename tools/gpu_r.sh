#!/bin/bash
# round-2 GPU call R (N GPUs, N = $1): strong scaling of the headline schedule with grouped pipelines; verified outputs
N=${1:-8}
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L | wc -l
run() {
  name=$1; shift
  env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 8 --warmup 3 --sweep none --no-cpu > gpurun_out/r02_r_n${N}_$name.json 2> gpurun_out/r02_r_n${N}_$name.err
  rc=$?
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02_r_n${N}_$name.json'))
    print('N=$N $name', 'ms', round(d['ms_per_step'],3), 'seq', round(d.get('create_proof_schedule_ms_no_ntt_overlap',0),3), 'msm_u', round(d['op_ms']['msm_uniform'],3), 'ver', d['verified']['msm'], d['verified']['msm_e2e'], d['verified']['all_ranks_ok'], 'e2e', round(d['e2e']['ms_per_step'],2), 'hostbuf', round(d['e2e_host_buffers']['ms_per_step'],2), 'c', d['msm_window_bits'])
except Exception as e:
    print('N=$N $name failed rc=$rc', e); print(open('gpurun_out/r02_r_n${N}_$name.err').read()[-1500:])
PY
}
run gdef H2B_MSM_GROUP=0
run g1 H2B_MSM_GROUP=1
