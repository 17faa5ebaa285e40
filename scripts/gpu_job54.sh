#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 300 python -m pytest tests/test_round2_fixes.py -q -m gpu -k "bucketed" --timeout 200 --timeout-method=thread --tb=short 2>&1 | tail -n 25 | cut -c1-250
timeout 500 python bench.py --no-cpu-baseline > gpurun_out/r2h_bench1.json 2> gpurun_out/r2h_bench1.err
tail -n 3 gpurun_out/r2h_bench1.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2h_bench1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')})
print('stream', d['small_batch'].get('stream_of_2_window_batches'))
print('2w', d['small_batch']['2_windows'])
PY
