#!/bin/bash
# round-2 GPU call D: ncu --set full of the batch-affine kernel (levels 3 and 1) at k = 19
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
H2B_AFF_LEVELS=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_batch_affine -s 1 -c 1 -f -o gpurun_out/r02_prof_ba_lv3 python tools/prof_ops.py 19 > gpurun_out/r02_d_ncu1.log 2>&1
echo "ncu1 rc=$?"
H2B_AFF_LEVELS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_batch_affine -s 1 -c 1 -f -o gpurun_out/r02_prof_ba_lv1 python tools/prof_ops.py 19 > gpurun_out/r02_d_ncu2.log 2>&1
echo "ncu2 rc=$?"
H2B_AFF_LEVELS=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_accumulate -s 1 -c 1 -f -o gpurun_out/r02_prof_acc_lv0 python tools/prof_ops.py 19 > gpurun_out/r02_d_ncu3.log 2>&1
echo "ncu3 rc=$?"
ls -la gpurun_out/*.ncu-rep
