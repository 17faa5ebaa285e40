#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
echo "== cluster on"; timeout 120 python scripts/epi_probe.py 5 2>&1 | tail -20
echo "== cluster off"; B3D_TMA_CLUSTER=0 timeout 120 python scripts/epi_probe.py 5 wide 2>&1 | tail -6
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_chain.py -q -m gpu --timeout 120 --timeout-method=thread --tb=short -x 2>&1 | tail -4 | cut -c1-300
