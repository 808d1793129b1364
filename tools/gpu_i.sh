#!/bin/bash
# round-2 GPU call I (2 GPUs): N>1 parity records — multi-rank pytest, bench at N=2 (self-verifying), fault injection, sanitizer
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_i_gpus.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -rA > gpurun_out/r02_i_pytest_multi.txt 2>&1
echo "multi pytest rc=$?"; tail -6 gpurun_out/r02_i_pytest_multi.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_i_bench_n2.json 2> gpurun_out/r02_i_bench_n2.err
echo "bench n2 rc=$?"; tail -3 gpurun_out/r02_i_bench_n2.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r02_i_bench_n2.json'))
    print('N2 ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'verified', {k:v for k,v in d['verified'].items() if k not in ('method','ntt')})
    for k,v in d['extra']['configs'].items(): print(k, v.get('create_proof_schedule_ms'), v.get('verified',{}).get('ok'), v.get('error'))
except Exception as e: print('bench parse failed', e)
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 1 --sweep none --no-cpu --inject-fault skip_allreduce > gpurun_out/r02_i_fault.json 2> gpurun_out/r02_i_fault.err
echo "fault injection rc=$? (expected non-zero)" | tee gpurun_out/r02_i_fault_rc.txt
grep "VERIFICATION FAILED" gpurun_out/r02_i_fault.err | head -2 | cut -c1-400
timeout 900 compute-sanitizer --tool racecheck --target-processes all python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r02_i_racecheck.txt 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|ERROR SUMMARY" gpurun_out/r02_i_racecheck.txt | tail -8
timeout 900 compute-sanitizer --tool memcheck --target-processes all python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r02_i_memcheck.txt 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r02_i_memcheck.txt | tail -8
