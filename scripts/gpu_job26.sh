#!/bin/bash
# final validation of the round: full GPU suite, smoke, default bench line, reference arm
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 2400 python -m pytest tests -q -m gpu --timeout 600 --timeout-method=thread --tb=short 2>&1 | grep -v "^    \|Warning\|warnings\|^$" | tail -12 | cut -c1-300
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 1500 python bench.py > gpurun_out/r2c_bench_default.json 2> gpurun_out/r2c_bench_default.err
tail -c 600 gpurun_out/r2c_bench_default.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2c_bench_default.json') if l.startswith('{')][-1])
for k in ('value','ms_per_step','gpu_launches','e2e','forward_only','roofline','roofline_step','cpu_baseline','batched_inference','fp32_mode','pose_model','small_batch','parity_timed_batch','clocks','peak_mem_gb'):
    print(k, json.dumps(d.get(k))[:700])
PY
timeout 900 python bench.py --impl reference > gpurun_out/r2c_bench_reference_arm.json 2> gpurun_out/r2c_bench_reference_arm.err
tail -c 400 gpurun_out/r2c_bench_reference_arm.json
