#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p /tmp/ncu
timeout 1200 ncu --set full --clock-control none -k regex:k_linear_tma -c 50 -o /tmp/ncu/r2_lean_classes python scripts/epi_probe.py 1 > gpurun_out/r2_lean_classes.log 2>&1
ncu -i /tmp/ncu/r2_lean_classes.ncu-rep --page raw --csv > gpurun_out/r2_lean_classes_raw.csv 2>/dev/null
ls -la /tmp/ncu gpurun_out/r2_lean_classes_raw.csv
gzip -f gpurun_out/r2_lean_classes_raw.csv
du -sh gpurun_out
