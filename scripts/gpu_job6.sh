#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
tail -2 gpurun_out/r2_bench_default.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2_bf16_launches.csv python bench.py --steps 1 --warmup 1 --profile-only > gpurun_out/r2_prof_bf16.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_fp32_launches.csv python bench.py --precision fp32 --scenes 8 --steps 1 --warmup 1 --profile-only > gpurun_out/r2_prof_fp32.log 2>&1
tail -1 gpurun_out/r2_prof_bf16.log; tail -1 gpurun_out/r2_prof_fp32.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_linear_tma -s 3 -c 1 -o gpurun_out/r2_roofline python scripts/roofline_kernel.py > gpurun_out/r2_roofline.log 2>&1
tail -2 gpurun_out/r2_roofline.log
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:"k_" -c 40 -o gpurun_out/r2_classes python scripts/class_probe2.py > gpurun_out/r2_classes.log 2>&1
tail -2 gpurun_out/r2_classes.log
timeout 600 python scripts/bench_loader.py --windows 360 --workers 8 --gpu > gpurun_out/r2_loader.txt 2>&1
cat gpurun_out/r2_loader.txt
