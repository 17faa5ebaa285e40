#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_ops.py tests/test_gpu_parity_bf16.py tests/test_gpu_chain.py -q -m gpu --timeout 300 --timeout-method=thread --tb=short 2>&1 | grep -v "^    \|Warning\|warnings\|^$" | tail -12 | cut -c1-400
for v in default edge_block; do
  if [ $v = edge_block ]; then export B3D_FEATURES=split_tc,window_knn,edge_block; fi
  timeout 400 python bench.py --steps 8 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2_bench_fe5_$v.json 2> gpurun_out/r2_bench_fe5_$v.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench_fe5_$v.json'))
print('$v', {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'fwd ms', d['forward_only']['ms_per_step'], 'e2e', d['e2e']['value'], 'mem', d['peak_mem_gb'])
PY
done
unset B3D_FEATURES
echo "== optime 32"; timeout 300 python scripts/optime.py 32 2>&1 | tail -42 | head -24
