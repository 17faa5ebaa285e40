#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
for g in 8 16; do echo "B3D_GR_ROWS=$g"; B3D_GR_ROWS=$g timeout 200 python scripts/seg_probe.py 32 2>&1 | grep gather_rows; done
B3D_GR_ROWS=16 timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "gather_rows" --timeout 200 --timeout-method=thread --tb=short 2>&1 | tail -n 2
