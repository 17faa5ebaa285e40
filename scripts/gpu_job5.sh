#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 600 python -m pytest tests/test_gpu_gather.py -x -q --timeout 200 --timeout-method=thread --tb=short > gpurun_out/r2_gather.log 2>&1
tail -25 gpurun_out/r2_gather.log | cut -c1-300
B3D_FEATURES=split_tc,window_knn,gather_tma timeout 400 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2_bench_gather.json 2> gpurun_out/r2_bench_gather.err
tail -3 gpurun_out/r2_bench_gather.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_gather.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['forward_only'], d['e2e']['value'])
PY
