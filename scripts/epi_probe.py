"""Epilogue probe for k_linear_tma: plain / ReLU-mask / gathered-addend variants at E = 1.96M rows
(run under ncu --set full to read stall reasons; prints CUDA-event times otherwise)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from batch3dmot_b200 import _lib as L, ops
dev = "cuda"
ops.set_precision("bf16")
E, N = 1958979, 64000
torch.manual_seed(0)
bf = torch.bfloat16
x256 = torch.randn(E, 256, device=dev).to(bf)
x128 = torch.randn(E, 128, device=dev).to(bf)
x64 = torch.randn(E, 64, device=dev).to(bf)
m256 = torch.randn(E, 256, device=dev).to(bf)
W = torch.randn(128, 256, device=dev) * 0.05          # Linear(256 -> 128)
Wm = torch.randn(192, 64, device=dev) * 0.05          # Linear(64 -> 192)
p = torch.randn(N, 192, device=dev).to(bf)
idx = torch.randint(0, N, (E,), device=dev).int().sort().values
cases = {
    "plain fwd 256->128": lambda: ops.linear_raw([(x256, None, None, 0)], W, None, E, L.ACT_RELU, tc=True, out_dtype=bf),
    "dgrad 128->256 + mask": lambda: ops.linear_raw([(x128, None, None, 0)], W, None, E, trans_w=True, out_mask=m256, tc=True, out_dtype=bf),
    "fwd 64->192 + gathered addend + relu": lambda: ops.linear_raw([(x64, None, None, 0)], Wm, None, E, L.ACT_RELU, tc=True, out_dtype=bf, adds=[(p, idx)]),
}
# the model's edge_update first layer: [e | att] (64 | 64) -> 256, two gathered addends (p_i[dst] streaming,
# p_j[src] local scatter), ReLU, sign bits out
W2 = torch.randn(256, 128, device=dev) * 0.05
pi, pj = torch.randn(N, 256, device=dev).to(bf), torch.randn(N, 256, device=dev).to(bf)
src_idx = (idx - torch.randint(50, 250, (E,), device=dev, dtype=torch.int32)).clamp_(0, N - 1)
bits = ops.new_relu_bits(E, 256, dev)
x64b = torch.randn(E, 64, device=dev).to(bf)
cases["fwd [64|64]->256 + 2 gathered addends + relu + bits"] = lambda: ops.linear_raw(
    [(x64, None, None, 0), (x64b, None, None, 0)], W2, None, E, L.ACT_RELU, tc=True, out_dtype=bf,
    adds=[(pi, idx), (pj, src_idx)], bits_out=bits)
# input-gradient layers of the backward pass with the producing layer's ReLU mask as sign bits
bits256 = ops.new_relu_bits(E, 256, dev); bits256.random_(-2**31, 2**31 - 1)
bits128 = ops.new_relu_bits(E, 128, dev); bits128.random_(-2**31, 2**31 - 1)
W64 = torch.randn(64, 128, device=dev) * 0.05          # Linear(128 -> 64): dgrad 64 -> 128
cases["dgrad 128->256 + sign-bit mask"] = lambda: ops.linear_raw([(x128, None, None, 0)], W, None, E, trans_w=True, mask_bits=bits256, tc=True, out_dtype=bf)
cases["dgrad 64->128 + sign-bit mask"] = lambda: ops.linear_raw([(x64, None, None, 0)], W64, None, E, trans_w=True, mask_bits=bits128, tc=True, out_dtype=bf)
cases["fwd 64->192 + gathered addend + relu + bits"] = lambda: ops.linear_raw([(x64, None, None, 0)], Wm, None, E, L.ACT_RELU, tc=True, out_dtype=bf, adds=[(p, idx)], bits_out=ops.new_relu_bits(E, 192, dev))
# att_edge_encoder first layer: 64 -> 512, two gathered addends
W5 = torch.randn(512, 64, device=dev) * 0.05
qi, qj = torch.randn(N, 512, device=dev).to(bf), torch.randn(N, 512, device=dev).to(bf)
bits512 = ops.new_relu_bits(E, 512, dev)
cases["fwd 64->512 + 2 gathered addends + relu + bits"] = lambda: ops.linear_raw(
    [(x64, None, None, 0)], W5, None, E, L.ACT_RELU, tc=True, out_dtype=bf, adds=[(qi, idx), (qj, src_idx)], bits_out=bits512)
bias128 = torch.randn(128, device=dev)
cases["plain fwd 256->128 + bias + bits"] = lambda: ops.linear_raw([(x256, None, None, 0)], W, bias128, E, L.ACT_RELU, tc=True, out_dtype=bf, bits_out=bits128)
cases["fwd 64->192 relu, no addend"] = lambda: ops.linear_raw([(x64, None, None, 0)], Wm, None, E, L.ACT_RELU, tc=True, out_dtype=bf)
pd = torch.randn(E, 192, device=dev).to(bf)
cases["fwd 64->192 relu + dense addend"] = lambda: ops.linear_raw([(x64, None, None, 0)], Wm, None, E, L.ACT_RELU, tc=True, out_dtype=bf, adds=[(pd, None)])
cases["fwd 64->256 relu, no addend"] = lambda: ops.linear_raw([(x64, None, None, 0)], W2[:, :64].contiguous(), None, E, L.ACT_RELU, tc=True, out_dtype=bf)
cases["fwd 64->128 relu, no addend"] = lambda: ops.linear_raw([(x64, None, None, 0)], W64.t().contiguous(), None, E, L.ACT_RELU, tc=True, out_dtype=bf)
# att_edge_encoder body: long-K layers (column blocks of a resident weight block)
x512 = torch.randn(E, 512, device=dev).to(bf)
x384 = torch.randn(E, 384, device=dev).to(bf)
W384 = torch.randn(384, 512, device=dev) * 0.05
W256 = torch.randn(256, 384, device=dev) * 0.05
b384 = torch.randn(384, device=dev)
cases["wide fwd 512->384 + bias + bits"] = lambda: ops.linear_raw([(x512, None, None, 0)], W384, b384, E, L.ACT_RELU, tc=True, out_dtype=bf, bits_out=ops.new_relu_bits(E, 384, dev))
cases["wide dgrad 384->512 + sign-bit mask"] = lambda: ops.linear_raw([(x384, None, None, 0)], W384, None, E, trans_w=True, mask_bits=bits512, tc=True, out_dtype=bf)
cases["wide fwd 384->256 + bits"] = lambda: ops.linear_raw([(x384, None, None, 0)], W256, None, E, L.ACT_RELU, tc=True, out_dtype=bf, bits_out=bits256)
out384 = torch.empty(E, 384, device=dev, dtype=bf)
cases["wide fwd 512->384 + bias (bench roofline launch, E rows)"] = lambda: ops.linear_raw([(x512, None, None, 0)], W384, b384, E, L.ACT_RELU, tc=True, out=out384)
Er = 489812
cases["wide fwd 512->384 + bias (bench roofline launch, 489,812 rows)"] = lambda: ops.linear_raw([(x512[:Er], None, None, 0)], W384, b384, Er, L.ACT_RELU, tc=True, out=out384[:Er])
only = sys.argv[2] if len(sys.argv) > 2 else None
if only:
    cases = {k: v for k, v in cases.items() if only in k}
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
for name, fn in cases.items():
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / reps * 1e3:.1f} us")
