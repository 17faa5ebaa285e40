#!/bin/bash
# usage: gpu_job_mg.sh N
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
tail -3 gpurun_out/r2_bench_${N}gpu.err
head -c 400 gpurun_out/r2_bench_${N}gpu.json
