#!/bin/bash
# round-2 GPU call AJ: ncu --set full of the bucket-reduction tail of the FINAL build (folded column weights)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:k_(rowcol_weights|weighted_final)" -c 2 -f -o gpurun_out/r02_prof_tail_final python tools/prof_ops.py 19 > gpurun_out/r02_aj_ncu.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/r02_prof_tail_final.ncu-rep
