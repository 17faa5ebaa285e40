#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 300 python scripts/torchprof.py 32 2>&1 | grep -v "^---" | cut -c1-230 | head -48
