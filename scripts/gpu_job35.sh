#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 1200 python -m pytest tests/test_gpu_plan_switches.py -q -m gpu --timeout 400 --timeout-method=thread --tb=short 2>&1 | grep -v "^    \|Warning\|warnings\|^$" | tail -14 | cut -c1-600
