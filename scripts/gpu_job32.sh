#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2c_fp32_launches.csv python bench.py --precision fp32 --scenes 8 --steps 1 --warmup 1 --profile-only > gpurun_out/r2c_prof_fp32.log 2>&1
python scripts/launch_summary.py gpurun_out/r2c_fp32_launches.csv 22
