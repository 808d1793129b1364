#!/bin/bash
# round-2 GPU call AD (N GPUs): the driver's own scaling command at N — full default bench (all sweep configs, proofs sharded)
N=${1:-8}
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
T0=$(date +%s)
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_ad_n${N}.json 2> gpurun_out/r02_ad_n${N}.err
echo "bench N=$N rc=$? wall $(( $(date +%s) - T0 )) s"
python - <<PY
import json
d=json.load(open('gpurun_out/r02_ad_n${N}.json'))
print('N=$N ms', round(d['ms_per_step'],3), 'placement', d.get('transform_placement'), 'ovl', d.get('create_proof_schedule_ms_ntt_overlap'), 'seq', d.get('create_proof_schedule_ms_no_ntt_overlap'), 'e2e', round(d['e2e']['ms_per_step'],3), 'hostbuf', round(d['e2e_host_buffers']['ms_per_step'],3), 'verified', d['verified']['all_ranks_ok'])
for k,v in d['extra']['configs'].items(): print(k, v.get('k'), round(v.get('create_proof_schedule_ms',0),3), (v.get('e2e_resident_proof') or {}).get('ms_per_proof'), v.get('verified',{}).get('ok'), v.get('error'))
PY
tail -3 gpurun_out/r02_ad_n${N}.err | cut -c1-300
