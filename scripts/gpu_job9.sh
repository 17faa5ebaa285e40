#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity_bf16.py -q -m gpu --timeout 300 --timeout-method=thread --tb=short 2>&1 | tail -3
echo "--- staged addends ON"; timeout 200 python scripts/gather_probe.py 32 2>&1 | grep -E "pre-projected|dense"
echo "--- staged addends OFF"; B3D_STAGE_ADDENDS=0 timeout 200 python scripts/gather_probe.py 32 2>&1 | grep -E "pre-projected"
