#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 2400 python -m pytest tests -q -m gpu --timeout 600 --timeout-method=thread --tb=short 2>&1 | grep -v "^    \|Warning\|warnings\|^$" | tail -30 | cut -c1-400
