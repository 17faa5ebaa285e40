"""Config-3 stress: frame-wise k-NN (warp top-k) + GAT aggregation throughput on synthetic graphs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from batch3dmot_b200 import ops, synth
from batch3dmot_b200.gat import GATConv
dev = "cuda"
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def timeit(fn, reps=10):
    for _ in range(2): fn()
    torch.cuda.synchronize(); ev0.record()
    for _ in range(reps): fn()
    ev1.record(); torch.cuda.synchronize()
    return ev0.elapsed_time(ev1) / reps * 1e-3
print("N, frame, D, k, skewed | knn us | Mquery/s | pair-dist GFLOP/s | gat us")
for N in (10000, 50000, 200000):
    for D in (48, 96):
        for k in (8, 16, 20):
            for skewed in (False, True):
                x, ptr = synth.knn_stress(1, N, frame=250, D=D, skewed=skewed)
                xd = x.to(dev); sizes = (ptr[1:] - ptr[:-1]).double()
                pairs = float((sizes * sizes).sum())
                t = timeit(lambda: ops.knn_frames(xd, ptr, k))
                idx = ops.knn_frames(xd, ptr, k)
                conv = GATConv(D, D).to(dev)
                with torch.no_grad():
                    tg = timeit(lambda: conv.forward_table(xd, idx))
                print(f"{N:7d} {250:4d} {D:3d} {k:3d} {int(skewed)} | {t*1e6:9.1f} | {N/t/1e6:8.2f} | {pairs*D*3/t/1e9:9.1f} | {tg*1e6:8.1f}")
