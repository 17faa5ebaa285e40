#!/bin/bash
# bring-up job: fused-chain tests first (bounded), then the whole GPU suite and the bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
export B3D_FEATURES=all
timeout 300 python -m pytest tests/test_gpu_chain.py -x -q --timeout 120 --timeout-method=thread > gpurun_out/r2_chain.log 2>&1
rc=$?
tail -25 gpurun_out/r2_chain.log
if [ $rc -ne 0 ]; then export B3D_FEATURES=split_tc,window_knn; echo "CHAIN TESTS FAILED rc=$rc -> continuing without the chain feature"; fi
timeout 1200 python -m pytest tests -m gpu -q --timeout 200 --timeout-method=thread --deselect tests/test_gpu_chain.py 2>&1 | tail -25 > gpurun_out/r2_t1.log
cat gpurun_out/r2_t1.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err
tail -5 gpurun_out/r2_bench1.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_ref1.json 2> gpurun_out/r2_ref1.err
tail -2 gpurun_out/r2_ref1.err
