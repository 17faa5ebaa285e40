#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 900 python -m pytest tests/test_gpu_split.py tests/test_gpu_models.py tests/test_gpu_ops.py -q -m gpu --timeout 300 --timeout-method=thread --tb=short 2>&1 | grep -v "^    \|Warning\|warnings\|^$" | tail -8 | cut -c1-300
timeout 600 python bench.py --precision fp32 --scenes 8 --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2_bench_fp32b.json 2> gpurun_out/r2_bench_fp32b.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench_fp32b.json'))
print('fp32', {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'fwd ms', d['forward_only']['ms_per_step'])
PY
