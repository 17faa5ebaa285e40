#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 600 python -m pytest tests/test_round2_fixes.py tests/test_gpu_parity_bf16.py tests/test_gpu_models.py -q -m gpu -x --timeout 600 --timeout-method=thread --tb=short 2>&1 | tail -n 12 | cut -c1-220
timeout 200 python scripts/small_batch_probe.py 2>&1 | tail -n 4
timeout 400 python bench.py --no-extras --no-cpu-baseline > gpurun_out/r2e_bench_ne3.json 2> gpurun_out/r2e_bench_ne3.err
tail -n 2 gpurun_out/r2e_bench_ne3.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2e_bench_ne3.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'fwd', d['forward_only']['ms_per_step'], 'e2e', d['e2e']['value'], 'roof', d['roofline']['frac'])
PY
