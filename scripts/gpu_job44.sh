#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
for f in "split_tc,window_knn,narrow_split,bf16_inputs" "split_tc,window_knn,narrow_split,bf16_inputs,chain" "split_tc,window_knn,narrow_split,bf16_inputs,edge_block" "split_tc,window_knn,bf16_inputs"; do
  B3D_FEATURES=$f timeout 200 python scripts/small_batch_probe.py 2>&1 | tail -n 4
done
