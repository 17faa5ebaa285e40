#!/bin/bash
# per-op timing of the current build at the bench workload (64 scenes)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 300 python scripts/optime.py 64 > gpurun_out/r2d_optime64.txt 2>&1
head -n 70 gpurun_out/r2d_optime64.txt | cut -c1-170
