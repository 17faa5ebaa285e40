#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity_bf16.py tests/test_gpu_chain.py tests/test_gpu_gather.py -q -m gpu --timeout 300 --timeout-method=thread --tb=short > gpurun_out/r2_t8.log 2>&1
tail -12 gpurun_out/r2_t8.log | cut -c1-300
echo "--- staged addends ON"; timeout 200 python scripts/gather_probe.py 32 2>&1 | grep -E "pre-projected|dense|x_j only"
echo "--- staged addends OFF"; B3D_STAGE_ADDENDS=0 timeout 200 python scripts/gather_probe.py 32 2>&1 | grep -E "pre-projected|x_j only"
timeout 400 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2_bench_stage.json 2> gpurun_out/r2_bench_stage.err
tail -3 gpurun_out/r2_bench_stage.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_stage.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['forward_only']['ms_per_step'], d['e2e']['value'])
PY
