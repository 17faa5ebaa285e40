#!/bin/bash
# 2-GPU bench line (weak headline + strong-scaling training key + sharded inference) and the capture_forward test
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 300 python -m pytest tests/test_round2_fixes.py -q -m gpu -k "inference_forward" --timeout 200 --timeout-method=thread --tb=short 2>&1 | tail -n 5 | cut -c1-250
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 > gpurun_out/r2e_bench_2gpu.json 2> gpurun_out/r2e_bench_2gpu.err
tail -n 3 gpurun_out/r2e_bench_2gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2e_bench_2gpu.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','n_gpus')})
print('strong', d['strong_scaling_training'])
print('infer', d['batched_inference']['scenes_per_s'])
PY
