#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
echo "== ca on"; timeout 120 python scripts/epi_probe.py 5 addend 2>&1 | tail -6
echo "== ca off"; B3D_ADD_CA=0 timeout 120 python scripts/epi_probe.py 5 addend 2>&1 | tail -6
for v in ca noca; do
  if [ $v = noca ]; then export B3D_ADD_CA=0; fi
  timeout 400 python bench.py --steps 8 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2_bench_fe10_$v.json 2> gpurun_out/r2_bench_fe10_$v.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench_fe10_$v.json'))
print('$v', {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'fwd ms', d['forward_only']['ms_per_step'], 'e2e', d['e2e']['value'])
PY
done
