"""Micro-benchmark of the bf16 segment sum (sorted and permuted) and the masked row gather at model shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from batch3dmot_b200 import ops
dev = "cuda"
host = bench.make_batch(0, int(sys.argv[1]) if len(sys.argv) > 1 else 32)
ei = host.edge_index.to(dev)
g = ops.Graph(ei, host.num_nodes)
E, N = g.E, g.N
bf = torch.bfloat16
def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for C in (192, 256):
    h = torch.randn(E, C, device=dev).to(bf)
    nb = E * C * 2
    for name, ni in (("by dst (sorted)", g.by_dst), ("by src (perm)", g.by_src)):
        us = t(lambda: ops.segment_sum_raw(h, ni))
        print(f"segment_sum bf16 C={C} {name}: {us:.1f} us  {nb / us / 1e3:.0f} GB/s")
    us = t(lambda: ops.segment_sum_raw(h, g.by_dst, out_dtype=bf))
    print(f"segment_sum bf16 C={C} by dst, bf16 out: {us:.1f} us  {nb / us / 1e3:.0f} GB/s")
dout = torch.randn(N, 192, device=dev)
bits = torch.randint(-2**31, 2**31 - 1, (6, E), device=dev, dtype=torch.int32)
for name, src in (("fp32 src", dout), ("bf16 src", dout.to(bf))):
    for ni, nn in ((g.by_dst, "dst"), (g.by_src, "src")):
        us = t(lambda: ops.gather_rows_raw(src, ni.idx, out_dtype=bf, relu_bits=bits))
        print(f"gather_rows {name} by {nn} + bits: {us:.1f} us  write {E * 384 / us / 1e3:.0f} GB/s")
