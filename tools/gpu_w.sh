#!/bin/bash
# round-2 GPU call W (2 GPUs): multi-rank tests, the new full-size grouped test, bench at N = 2 with the sweep proofs sharded
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_fullsize.py -m gpu -q -k "multi or single_process or grouped" > gpurun_out/r02_w_pytest.txt 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r02_w_pytest.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 8 --warmup 3 --sweep 1,2 --no-cpu > gpurun_out/r02_w_bench_n2.json 2> gpurun_out/r02_w_bench_n2.err
echo "bench n2 rc=$?"; tail -3 gpurun_out/r02_w_bench_n2.err | cut -c1-600
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_w_bench_n2.json'))
print('N2 ms', round(d['ms_per_step'],3), 'ovl', d.get('create_proof_schedule_ms_ntt_overlap'), 'seq', d.get('create_proof_schedule_ms_no_ntt_overlap'), 'e2e', round(d['e2e']['ms_per_step'],3), 'verified', {k:v for k,v in d['verified'].items() if k not in ('method','ntt')})
for k,v in d['extra']['configs'].items(): print(k, v.get('create_proof_schedule_ms'), v.get('e2e_resident_proof'), v.get('verified',{}).get('ok'), v.get('error'))
PY
