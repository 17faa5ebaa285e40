#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
echo "== default"; timeout 300 python scripts/torchprof.py 32 2>&1 | grep -v "^---" | cut -c1-200 | grep "Self CUDA time total\|k_add_n\|k_linear_tma<0, true>\|k_linear_tma<1, true>\|k_wgrad_tma\|aten::copy_\|aten::contiguous\|aten::cat \|Memcpy\|elementwise_kernel" | head -14
echo "== edge_block"; B3D_FEATURES=split_tc,window_knn,narrow_split,bf16_inputs,edge_block timeout 300 python scripts/torchprof.py 32 2>&1 | grep -v "^---" | cut -c1-200 | grep "Self CUDA time total\|k_add_n\|k_linear_tma<0, true>\|k_linear_tma<1, true>\|k_linear_tma<0, false>\|k_wgrad_tma\|aten::copy_\|aten::contiguous\|aten::cat \|Memcpy\|elementwise_kernel\|_MPEdgeBlockG" | head -16
