#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
export B3D_FEATURES=all
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread --tb=short > gpurun_out/r2_t2.log 2>&1
tail -40 gpurun_out/r2_t2.log
timeout 300 python scripts/chain_probe.py 32 > gpurun_out/r2_chain_probe.txt 2>&1
cat gpurun_out/r2_chain_probe.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_launches_chain.csv python bench.py --steps 1 --warmup 1 --profile-only > gpurun_out/r2_prof_chain.log 2>&1
tail -2 gpurun_out/r2_prof_chain.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_chain -s 10 -c 4 -o gpurun_out/r2_chain_full python scripts/chain_probe.py 8 > gpurun_out/r2_ncu_chain.log 2>&1
tail -3 gpurun_out/r2_ncu_chain.log
B3D_FEATURES=none timeout 300 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2_bench_nofeat.json 2> gpurun_out/r2_bench_nofeat.err
head -c 300 gpurun_out/r2_bench_nofeat.json
