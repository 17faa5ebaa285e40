#!/bin/bash
# last verification of the round: smoke(), the full -m gpu suite, the default bench line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) 2>&1 | tail -n 5
( time timeout 1200 python -m pytest tests -q -m gpu --timeout 600 --timeout-method=thread --tb=short ) > gpurun_out/r2i_gputests.log 2>&1
tail -n 6 gpurun_out/r2i_gputests.log | cut -c1-200
( time timeout 600 python bench.py ) > gpurun_out/r2i_bench_default.json 2> gpurun_out/r2i_bench_default.err
tail -n 4 gpurun_out/r2i_bench_default.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2i_bench_default.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'fwd', d['forward_only']['ms_per_step'], 'e2e', d['e2e']['value'], 'roof', d['roofline']['frac'], d['clocks'])
print(d['small_batch']['2_windows'])
PY
