#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
( timeout 1200 python -m pytest tests -q -m gpu -x --timeout 600 --timeout-method=thread --tb=short ) > gpurun_out/r2e_gputests.log 2>&1
tail -n 6 gpurun_out/r2e_gputests.log | cut -c1-200
timeout 400 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r2e_bench_ne.json 2> gpurun_out/r2e_bench_ne.err
tail -n 3 gpurun_out/r2e_bench_ne.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2e_bench_ne.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'fwd', d['forward_only']['ms_per_step'], 'e2e', d['e2e']['value'], 'roof', d['roofline']['frac'])
PY
