#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
export B3D_FEATURES=split_tc,window_knn
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread --tb=short > gpurun_out/r2_t4.log 2>&1
tail -30 gpurun_out/r2_t4.log | cut -c1-400
timeout 200 python -c "
import __graft_entry__ as g
g.smoke()
" 2>&1 | tail -3
