#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
for v in 1 0; do
B3D_STAGE_ADDENDS=$v timeout 400 python bench.py --steps 8 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2_bench_stage$v.json 2> gpurun_out/r2_bench_stage$v.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench_stage$v.json'))
print('stage=$v', {k:d[k] for k in ('value','ms_per_step')}, 'fwd ms', d['forward_only']['ms_per_step'], 'e2e', d['e2e']['value'])
PY
done
