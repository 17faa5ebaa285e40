#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
export B3D_FEATURES=all
timeout 300 python -m pytest tests/test_gpu_chain.py -x -q --timeout 120 --timeout-method=thread --tb=short > gpurun_out/r2_chain3.log 2>&1
tail -15 gpurun_out/r2_chain3.log
timeout 300 python scripts/chain_probe.py 32 > gpurun_out/r2_chain_probe3.txt 2>&1
cat gpurun_out/r2_chain_probe3.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_chain -s 10 -c 2 -o gpurun_out/r2_chain_full3 python scripts/chain_probe.py 8 > gpurun_out/r2_ncu_chain3.log 2>&1
tail -2 gpurun_out/r2_ncu_chain3.log
