#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_linear_tma -c 4 -o gpurun_out/r2_wide python scripts/epi_probe.py 1 "wide fwd 512" > gpurun_out/r2_wide_ncu.log 2>&1
tail -2 gpurun_out/r2_wide_ncu.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_linear_tma -c 4 -o gpurun_out/r2_add2 python scripts/epi_probe.py 1 "[64|64]" > gpurun_out/r2_add2_ncu.log 2>&1
tail -2 gpurun_out/r2_add2_ncu.log
