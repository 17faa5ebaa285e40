#!/bin/bash
# N-GPU bench line as the driver launches it: usage gpu_job49.sh N
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N > gpurun_out/r2f_bench_${N}gpu.json 2> gpurun_out/r2f_bench_${N}gpu.err
tail -n 3 gpurun_out/r2f_bench_${N}gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2f_bench_${N}gpu.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','n_gpus')}, 'e2e', d['e2e']['value'])
print('strong', d['strong_scaling_training'])
print('small dp', d['small_batch_data_parallel'])
print('infer', d['batched_inference']['scenes_per_s'])
PY
