#!/bin/bash
# full -m gpu suite on the current build + the default bench line (as the driver runs it)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
( time timeout 1200 python -m pytest tests -q -m gpu -x --timeout 600 --timeout-method=thread --tb=short --durations=15 ) > gpurun_out/r2d_gputests.log 2>&1
tail -n 25 gpurun_out/r2d_gputests.log | cut -c1-200
( time timeout 600 python bench.py ) > gpurun_out/r2d_bench_default.json 2> gpurun_out/r2d_bench_default.err
tail -n 4 gpurun_out/r2d_bench_default.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2d_bench_default.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'fwd', d['forward_only'], 'e2e', d['e2e']['value'], 'roof', d['roofline'])
PY
