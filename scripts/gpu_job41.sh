#!/bin/bash
# 2-GPU bench line (weak headline + strong-scaling training key + small-batch DP + sharded inference)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 > gpurun_out/r2e_bench_2gpu.json 2> gpurun_out/r2e_bench_2gpu.err
tail -n 3 gpurun_out/r2e_bench_2gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2e_bench_2gpu.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','n_gpus')})
print('strong', d['strong_scaling_training'])
print('small dp', d['small_batch_data_parallel'])
print('small', d['small_batch'])
print('infer', d['batched_inference']['scenes_per_s'])
PY
