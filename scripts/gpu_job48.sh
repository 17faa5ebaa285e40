#!/bin/bash
# final evidence of the session, 1 GPU: full -m gpu suite, the default bench line, the reference arm, the launch list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
( time timeout 1200 python -m pytest tests -q -m gpu --timeout 600 --timeout-method=thread --tb=short ) > gpurun_out/r2f_gputests.log 2>&1
tail -n 6 gpurun_out/r2f_gputests.log | cut -c1-200
( time timeout 600 python bench.py ) > gpurun_out/r2f_bench_default.json 2> gpurun_out/r2f_bench_default.err
tail -n 4 gpurun_out/r2f_bench_default.err
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/r2f_bench_reference_arm.json 2> gpurun_out/r2f_bench_reference_arm.err
tail -n 4 gpurun_out/r2f_bench_reference_arm.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2f_bf16_launches.csv python bench.py --steps 1 --warmup 1 --profile-only > gpurun_out/r2f_prof_bf16.log 2>&1
tail -n 2 gpurun_out/r2f_prof_bf16.log
python scripts/launch_summary.py gpurun_out/r2f_bf16_launches.csv 30
python - <<PY
import json
d=json.loads(open('gpurun_out/r2f_bench_default.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'fwd', d['forward_only']['ms_per_step'], 'e2e', d['e2e']['value'], 'roof', d['roofline']['frac'])
r=json.loads(open('gpurun_out/r2f_bench_reference_arm.json').read().strip().splitlines()[-1])
print('ref', r['value'], r['cpu_baseline']['kind'])
PY
