#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 200 python -m pytest tests/test_round2_fixes.py -q -m gpu -k "bucketed" --timeout 150 --timeout-method=thread --tb=short 2>&1 | tail -n 12 | cut -c1-300
