#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 300 python scripts/host_profile.py 2>&1 | head -n 22 | cut -c1-160
timeout 300 python -m pytest tests/test_graph_io.py tests/test_round2_fixes.py tests/test_tracking.py -q -m gpu --timeout 200 --timeout-method=thread --tb=short 2>&1 | tail -n 2
