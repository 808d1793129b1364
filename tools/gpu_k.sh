#!/bin/bash
# round-2 GPU call K (2 GPUs): everything added since call I — full GPU suite (incl. device group, Assigned ingestion,
# compressed points, lookup orders, C++ mirror group section), single-process bench vs one-process-per-GPU bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02_k_pytest_all.txt 2>&1
echo "all pytest rc=$?"; tail -25 gpurun_out/r02_k_pytest_all.txt
timeout 900 python bench.py --gpus 2 --single-process --steps 5 --warmup 2 > gpurun_out/r02_k_bench_sp2.json 2> gpurun_out/r02_k_bench_sp2.err
echo "single-process N=2 rc=$?"; tail -3 gpurun_out/r02_k_bench_sp2.err
timeout 900 python bench.py --gpus 1 --single-process --steps 5 --warmup 2 > gpurun_out/r02_k_bench_sp1.json 2> gpurun_out/r02_k_bench_sp1.err
echo "single-process N=1 rc=$?"; tail -3 gpurun_out/r02_k_bench_sp1.err
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 5 --warmup 2 --sweep none --no-cpu > gpurun_out/r02_k_bench_mp2.json 2> gpurun_out/r02_k_bench_mp2.err
echo "multi-process N=2 rc=$?"
python - <<'PY'
import json
for f in ('sp2','sp1','mp2'):
    try:
        d=json.load(open(f'gpurun_out/r02_k_bench_{f}.json'))
        print(f, 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'host-buffers', d.get('e2e_host_buffers',{}).get('ms_per_step_sequential_calls'), d.get('e2e_host_buffers',{}).get('ms_per_step_overlapped'), 'verified', d['verified'].get('msm_e2e'), d['verified'].get('of'))
    except Exception as e: print(f, 'parse failed', e)
PY
