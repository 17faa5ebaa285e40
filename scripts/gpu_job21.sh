#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_chain.py -q -m gpu --timeout 300 --timeout-method=thread --tb=short 2>&1 | tail -3 | cut -c1-300
echo "== wide on"; timeout 200 python scripts/epi_probe.py 5 wide 2>&1 | tail -3
echo "== wide off"; B3D_TMA_WIDE=0 timeout 200 python scripts/epi_probe.py 5 wide 2>&1 | tail -3
for v in wide nowide; do
  if [ $v = nowide ]; then export B3D_TMA_WIDE=0; fi
  timeout 400 python bench.py --steps 8 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2_bench_fe6_$v.json 2> gpurun_out/r2_bench_fe6_$v.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench_fe6_$v.json'))
print('$v', {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'fwd ms', d['forward_only']['ms_per_step'], 'e2e', d['e2e']['value'], 'mem', d['peak_mem_gb'])
PY
done
