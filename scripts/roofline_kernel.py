"""Runs only the dominant-kernel launch that bench.py quotes in `roofline` (k_linear_tma on
att_edge_encoder layer 2, [E,512] bf16 -> 384, ReLU, bf16 out, E = 489,812) so that
`ncu --set full -k regex:k_linear_tma` can capture its DRAM traffic."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from batch3dmot_b200 import ops
ops.set_precision("bf16")
E = 489812
torch.manual_seed(0)
W = torch.randn(384, 512, device="cuda") * 0.05; b = torch.randn(384, device="cuda")
h = torch.randn(E, 512, device="cuda").to(torch.bfloat16)
out = torch.empty(E, 384, device="cuda", dtype=torch.bfloat16)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3):
    ops.linear_raw([(h, None, None, 0)], W, b, E, 1, out=out, tc=True)
torch.cuda.synchronize(); ev0.record()
for _ in range(10):
    ops.linear_raw([(h, None, None, 0)], W, b, E, 1, out=out, tc=True)
ev1.record(); torch.cuda.synchronize()
t = ev0.elapsed_time(ev1) / 10 * 1e-3
print(f"k_linear_tma att2: {t*1e6:.1f} us  {2.0*E*512*384/t/1e12:.1f} TFLOP/s  alg {E*(512+384)*2/t/1e9:.0f} GB/s")
