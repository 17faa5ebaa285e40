#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
export B3D_STAGE_SPLIT=0
echo "== probe"; timeout 200 python scripts/epi_probe.py 5 2>&1 | tail -14
echo "== probe no staging"; B3D_STAGE_ADDENDS=0 timeout 200 python scripts/epi_probe.py 5 "64->192" 2>&1 | tail -6
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_linear_tma -c 6 -o gpurun_out/r2_epi python scripts/epi_probe.py 1 "64->192" > gpurun_out/r2_epi_ncu.log 2>&1
tail -3 gpurun_out/r2_epi_ncu.log
