#!/bin/bash
# round-2 GPU call AB: timeline of one resident PROOF (every launch with stream, ready and finished time)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python tools/timeline.py --prover --out gpurun_out/r02_ab_timeline_prover.csv > gpurun_out/r02_ab_timeline_prover.txt 2>&1; echo "rc=$?"; head -40 gpurun_out/r02_ab_timeline_prover.txt
