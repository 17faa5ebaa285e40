#!/bin/bash
# capture_forward test + bench line with the new keys (1 GPU)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 300 python -m pytest tests/test_round2_fixes.py -q -m gpu -k "inference_forward" --timeout 200 --timeout-method=thread --tb=short 2>&1 | tail -n 15 | cut -c1-250
timeout 500 python bench.py --no-cpu-baseline > gpurun_out/r2e_bench1.json 2> gpurun_out/r2e_bench1.err
tail -n 3 gpurun_out/r2e_bench1.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2e_bench1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')})
print('pose', d['pose_model']['1_scene'])
print('strong', d['strong_scaling_training'])
PY
