"""Bring-up diagnostics for the tcgen05 kernels: prints max relative error of b3d_linear_tc /
b3d_wgrad_tc against a float64 reference on bf16-rounded operands, for both descriptor
conventions (b3d_tc_debug bit0). Run on the GPU box."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from batch3dmot_b200 import _lib as L, ops

dev = "cuda"
bf = lambda t: t.to(torch.bfloat16).to(torch.float64)


def rel(a, b):
    return float((a.double().cpu() - b.cpu()).abs().max() / b.abs().max())


def run_linear(M, widths, n_out, gather, trans=False):
    torch.manual_seed(M + n_out)
    N = 777
    K = sum(widths)
    xs = [torch.randn(N if (gather and s % 2 == 0) else M, w) for s, w in enumerate(widths)]
    idx = torch.randint(0, N, (len(widths), M))
    cat = torch.cat([x[idx[s]] if x.size(0) == N else x for s, x in enumerate(xs)], 1)
    items = [(x.to(dev), idx[s].int().to(dev) if x.size(0) == N else None, None, 0) for s, x in enumerate(xs)]
    if trans:
        W = torch.randn(K, n_out) * 0.1      # stored [K_fwd_out.., ] : Y = A @ W
        ref = bf(cat) @ bf(W)
    else:
        W = torch.randn(n_out, K) * 0.1
        ref = bf(cat) @ bf(W).t()
    b = torch.randn(n_out)
    ref = torch.relu(ref + b.double())
    y = ops.linear_raw(items, W.to(dev), b.to(dev), M, L.ACT_RELU, trans_w=trans)
    torch.cuda.synchronize()
    return rel(y, ref)


def run_wgrad(M, widths, n_out):
    torch.manual_seed(M)
    N = 333
    xs = [torch.randn(N if s % 2 == 0 else M, w) for s, w in enumerate(widths)]
    idx = torch.randint(0, N, (len(widths), M))
    cat = torch.cat([x[idx[s]] if x.size(0) == N else x for s, x in enumerate(xs)], 1)
    dy, y = torch.randn(M, n_out), torch.randn(M, n_out)
    items = [(x.to(dev), idx[s].int().to(dev) if x.size(0) == N else None, None, 0) for s, x in enumerate(xs)]
    dW, db = ops.wgrad_raw((dy.to(dev), None, y.to(dev), L.MASK_RELU), items, M, n_out, sum(widths))
    torch.cuda.synchronize()
    dym = dy * (y > 0)
    return rel(dW, bf(dym).t() @ bf(cat)), rel(db, bf(dym).sum(0))


ops.set_precision("bf16")
for dbg in (0,):
    L.lib().b3d_tc_debug(dbg)
    print(f"--- dbg={dbg}")
    for args in [(1024, (64,), 64, False), (3000, (48, 48, 32), 96, True), (5000, (96, 96, 64, 64), 256, True),
                 (2000, (64, 128, 96, 64, 128, 96, 64), 512, True), (1500, (256,), 128, False), (700, (128,), 48, False)]:
        try:
            print("linear", args, "rel err %.3e" % run_linear(*args))
        except Exception as e:
            print("linear", args, "FAILED", e)
    print("linear trans", "rel err %.3e" % run_linear(4000, (256,), 320, False, trans=True))
    for args in [(4096, (64,), 64), (3000, (48, 48, 32), 96), (70000, (96, 96, 64, 64), 256), (9000, (96, 64, 96), 192),
                 (5000, (128,), 64), (20000, (640,), 512)]:
        try:
            print("wgrad", args, "rel err dW %.3e db %.3e" % run_wgrad(*args))
        except Exception as e:
            print("wgrad", args, "FAILED", e)
