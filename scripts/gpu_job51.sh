#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 300 python scripts/host_profile.py > gpurun_out/r2f_host_profile.txt 2>&1
head -n 75 gpurun_out/r2f_host_profile.txt | cut -c1-200
