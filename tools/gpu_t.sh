#!/bin/bash
# round-2 GPU call T: full GPU suite + ncu launch list of a bench step + ncu --set full of the round-2 kernels
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r02_t_pytest_all.txt 2>&1
echo "all pytest rc=$?"; tail -4 gpurun_out/r02_t_pytest_all.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 1 --sweep none --no-cpu > gpurun_out/r02_t_launches_bench.json 2> gpurun_out/r02_t_launches_bench.err
echo "launch list rc=$?"; wc -l gpurun_out/r02_launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_accumulate -s 38 -c 1 -f -o gpurun_out/r02_prof_acc_group4 python tools/prof_ops.py 19 > gpurun_out/r02_t_ncu1.log 2>&1
echo "ncu acc rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_(rowcol_sums|rowcol_weights|weighted_final|collect|digits)" -c 7 -f -o gpurun_out/r02_prof_tail python tools/prof_ops.py 19 > gpurun_out/r02_t_ncu2.log 2>&1
echo "ncu tail rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ntt_pass -c 4 -f -o gpurun_out/r02_prof_ntt python tools/prof_ops.py 19 > gpurun_out/r02_t_ncu3.log 2>&1
echo "ncu ntt rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_radix_sort -c 2 -f -o gpurun_out/r02_prof_sort python tools/prof_quotient.py 19 --quick > gpurun_out/r02_t_ncu4.log 2>&1
echo "ncu sort rc=$?"; tail -3 gpurun_out/r02_t_ncu4.log
ls -la gpurun_out/r02_prof_*.ncu-rep
