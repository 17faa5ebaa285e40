#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity_bf16.py tests/test_gpu_models.py tests/test_gpu_real_encoders.py -q -m gpu --timeout 300 --timeout-method=thread --tb=short 2>&1 | grep -v "^    \|Warning\|warnings\|^$" | tail -8 | cut -c1-300
for v in default noconv; do
  if [ $v = noconv ]; then export B3D_FEATURES=split_tc,window_knn,narrow_split; fi
  timeout 400 python bench.py --steps 8 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2_bench_fe9_$v.json 2> gpurun_out/r2_bench_fe9_$v.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench_fe9_$v.json'))
print('$v', {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'fwd ms', d['forward_only']['ms_per_step'], 'e2e', d['e2e']['value'], 'mem', d['peak_mem_gb'])
print(json.dumps(d['roofline_step'])[:900])
PY
done
