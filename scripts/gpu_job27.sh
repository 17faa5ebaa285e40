#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "narrow_chains_split" --timeout 300 --timeout-method=thread --tb=short 2>&1 | grep "Error\|assert " | cut -c1-200 | head
python - <<'PY'
import torch, sys
sys.path.insert(0,'.')
sys.path.insert(0,'tests')
from batch3dmot_b200 import ops, _lib as L
import test_gpu_ops as T
DEV='cuda'
torch.manual_seed(11)
M=20000
ops.set_precision("bf16")
for dims, sig in (((64, 32, 16, 8, 1), True), ((4, 16, 32, 64), False)):
    seq = T._torch_chain(dims, sig).to(DEV)
    x = torch.randn(M, dims[0], device=DEV)
    x = x.to(torch.bfloat16) if dims[0] == 64 else x.double()
    xr = x.float().requires_grad_(True)
    y_ref = seq(xr)
    gy = torch.randn_like(y_ref)
    gref = torch.autograd.grad(y_ref, [xr] + list(seq.parameters()), gy)
    xin = x.clone().requires_grad_(dims[0] == 64)
    y = ops.run_mlp(seq, [(ops.edge_attr_rows(xin) if dims[0] == 4 else xin, None)], final_act="sigmoid" if sig else None, out_dtype=torch.float32 if sig else torch.bfloat16)
    print(dims, 'fwd rel', T.rel_err(y, y_ref))
    got = torch.autograd.grad(y, ([xin] if dims[0] == 64 else []) + list(seq.parameters()), gy.to(y.dtype))
    for i,(g, r) in enumerate(zip(got, gref[0 if dims[0] == 64 else 1:])):
        print('  grad', i, tuple(g.shape), float((g.double() - r.double()).norm() / r.double().norm()))
PY
