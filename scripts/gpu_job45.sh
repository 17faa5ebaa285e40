#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2e_small_launches.csv python scripts/small_batch_step.py 2 > gpurun_out/r2e_small_prof.log 2>&1
tail -n 2 gpurun_out/r2e_small_prof.log
python scripts/launch_summary.py gpurun_out/r2e_small_launches.csv 30
