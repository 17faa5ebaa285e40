import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from batch3dmot_b200 import _lib as L, ops
dev = "cuda"
ops.set_precision("bf16")
torch.manual_seed(0)
def rel(a, b): return float((a.double() - b.double()).abs().max() / b.double().abs().max())
for M in (61561, 16000):
    for (k, n) in [(512, 384), (384, 256), (256, 128), (128, 64), (64, 512), (192, 128), (256, 192), (96, 256), (288, 512), (128, 96)]:
        for trans in (False, True):
            x = torch.randn(M, n if trans else k, device=dev).to(torch.bfloat16)
            W = torch.randn(n, k, device=dev) * 0.1
            hm = torch.randn(M, k if trans else n, device=dev).to(torch.bfloat16)
            for od in (torch.bfloat16, torch.float32):
                ops._USE_TMA = True
                y1 = ops.linear_raw([(x, None, None, 0)], W, None, M, 0, trans_w=trans, out_mask=hm, tc=True, out_dtype=od)
                ops._USE_TMA = False
                y0 = ops.linear_raw([(x, None, None, 0)], W, None, M, 0, trans_w=trans, out_mask=hm, tc=True, out_dtype=od)
                torch.cuda.synchronize()
                r = rel(y1, y0)
                if r > 1e-3: print("MISMATCH", M, k, n, trans, od, r)
print("done")
