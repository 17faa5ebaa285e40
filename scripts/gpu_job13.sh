#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 600 python -m pytest tests/test_gpu_tc.py -q -m gpu --timeout 200 --timeout-method=thread --tb=short -x 2>&1 | tail -3 | cut -c1-300
echo "== probe default"; timeout 200 python scripts/epi_probe.py 5 2>&1 | tail -12
echo "== probe SPLIT=0"; B3D_STAGE_SPLIT=0 timeout 200 python scripts/epi_probe.py 5 2>&1 | tail -12
timeout 900 python -m pytest tests/test_gpu_parity_bf16.py -q -m gpu --timeout 300 --timeout-method=thread --tb=short 2>&1 | grep -v "^    \|Warning\|warnings" | tail -40 | cut -c1-400
for v in default nosplit; do
  if [ $v = nosplit ]; then export B3D_STAGE_SPLIT=0; fi
  timeout 400 python bench.py --steps 8 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2_bench_fe_$v.json 2> gpurun_out/r2_bench_fe_$v.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench_fe_$v.json'))
print('$v', {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'fwd ms', d['forward_only']['ms_per_step'], 'e2e', d['e2e']['value'], 'mem', d['peak_mem_gb'])
PY
done
