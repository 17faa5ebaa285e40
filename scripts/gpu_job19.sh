#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2b_bf16_launches.csv python bench.py --steps 1 --warmup 1 --profile-only > gpurun_out/r2b_prof_bf16.log 2>&1
python scripts/launch_summary.py gpurun_out/r2b_bf16_launches.csv 30
