"""The opt-in kernel plans of b3d_linear_tma (environment switches read once per process by libb3d.so, see
INTEGRATION.md) stay correct: every switch runs the same layer set in a fresh interpreter and compares with a float64
reference on the bf16-rounded operands."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import torch
from batch3dmot_b200 import _lib as L, ops
ops.set_precision("bf16")
dev = "cuda"
bf = lambda t: t.to(torch.bfloat16).to(torch.float64)
torch.manual_seed(0)
M, N = 40000, 700
worst = 0.0
for K, n_out, nadd, bits in ((512, 384, 0, True), (384, 512, 0, False), (128, 256, 2, True), (64, 512, 2, True),
                             (64, 192, 1, False), (256, 128, 0, True)):
    x = torch.randn(M, K).to(torch.bfloat16)
    W = torch.randn(n_out, K) * 0.1
    b = torch.randn(n_out) if nadd == 0 else None
    adds, ref = [], bf(x) @ bf(W).t() + (b.double() if b is not None else 0)
    for t in range(nadd):
        p = torch.randn(N, n_out).to(torch.bfloat16)
        i = torch.randint(0, N, (M,))
        if t == 0:
            i = i.sort().values
        adds.append((p.to(dev), i.int().to(dev)))
        ref = ref + p.double()[i]
    ref = torch.relu(ref)
    bt = ops.new_relu_bits(M, n_out, dev) if bits else None
    y = ops.linear_raw([(x.to(dev), None, None, 0)], W.to(dev), b.to(dev) if b is not None else None, M, L.ACT_RELU, tc=True,
                       out_dtype=torch.bfloat16, adds=adds or None, bits_out=bt)
    err = float((y.double().cpu() - ref).abs().max() / ref.abs().max())
    worst = max(worst, err)
    assert err < 1e-2, (K, n_out, nadd, err)
    if bits:
        un = ((bt.t().unsqueeze(2) >> torch.arange(32, device=dev, dtype=torch.int32)) & 1).reshape(M, -1)[:, :n_out].bool()
        assert torch.equal(un, y > 0), (K, n_out)
        gz = torch.randn(M, 64).to(torch.bfloat16).to(dev)
        Wd = torch.randn(64, n_out, device=dev) * 0.1
        d1 = ops.linear_raw([(gz, None, None, 0)], Wd, None, M, trans_w=True, tc=True, mask_bits=bt, out_dtype=torch.bfloat16)
        d2 = ops.linear_raw([(gz, None, None, 0)], Wd, None, M, trans_w=True, tc=True, out_mask=y, out_dtype=torch.bfloat16)
        assert torch.equal(d1, d2), (K, n_out)
torch.cuda.synchronize()
print("OK", worst)
'''


@pytest.mark.parametrize("env", [{}, {"B3D_TMA_CLUSTER": "1"}, {"B3D_TMA_WIDE": "1"}, {"B3D_ADD_CA": "1"},
                                 {"B3D_STAGE_SPLIT": "0"}, {"B3D_TMA_DEEP": "0"}, {"B3D_FAST_EPI": "0"},
                                 {"B3D_STAGE_ADDENDS": "0"}, {"B3D_STAGE_PLAN": "1,1,4"}],
                         ids=lambda e: ",".join(f"{k}={v}" for k, v in e.items()) or "default")
def test_linear_tma_plan_switch(env):
    full = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""), **env)
    r = subprocess.run([sys.executable, "-c", SCRIPT], cwd=ROOT, env=full, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "OK" in r.stdout, (r.stdout[-2000:], r.stderr[-3000:])
