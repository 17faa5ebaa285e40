#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 300 python scripts/torchprof.py 64 > gpurun_out/r2e_torchprof64.txt 2>&1
head -n 60 gpurun_out/r2e_torchprof64.txt | cut -c1-260
