"""Per-layer micro-benchmark of the dense kernels on the multimodal message-passing shapes.
Usage: python scripts/bench_layers.py [bf16|fp32] [E] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from batch3dmot_b200 import _lib as L, ops, synth

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
scenes = int(sys.argv[2]) if len(sys.argv) > 2 else 8
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
dev = "cuda"
ops.set_precision(prec)
c = synth.collate([synth.scene_graph(seed=5621 + i) for i in range(scenes)])
G = ops.Graph(c.edge_index.to(dev), c.num_nodes)
E, N = G.E, G.N
x = torch.randn(N, 96, device=dev); x0 = torch.randn(N, 96, device=dev)
e = torch.randn(E, 64, device=dev); att = torch.randn(E, 64, device=dev)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timeit(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(reps):
        fn()
    ev1.record()
    torch.cuda.synchronize()
    return ev0.elapsed_time(ev1) / reps * 1e-3


def layer(name, items, n_out, act=1):
    K = sum(t.size(1) for t, _, _, _ in items)
    W = torch.randn(n_out, K, device=dev) * 0.05
    b = torch.randn(n_out, device=dev)
    M = items[0][1].numel() if items[0][1] is not None else items[0][0].size(0)
    out = torch.empty(M, n_out, device=dev)
    t = timeit(lambda: ops.linear_raw(items, W, b, M, act, out=out))
    fl = 2.0 * M * K * n_out
    by = 4.0 * M * (K + n_out)
    print(f"fwd   {name:28s} M={M:7d} K={K:4d} N={n_out:4d}  {t*1e6:9.1f} us  {fl/t/1e12:7.1f} TFLOP/s  {by/t/1e9:7.0f} GB/s(alg)")
    dy = torch.randn(M, n_out, device=dev)
    dA = torch.empty(M, K, device=dev)
    hm = torch.randn(M, K, device=dev)
    t = timeit(lambda: ops.linear_raw([(dy, None, None, 0)], W, None, M, trans_w=True, out=dA, out_mask=hm))
    print(f"dgrad {name:28s} {'':36s}{t*1e6:9.1f} us  {fl/t/1e12:7.1f} TFLOP/s")
    dW = torch.empty(n_out, K, device=dev); db = torch.empty(n_out, device=dev)
    t = timeit(lambda: ops.wgrad_raw((dy, None, None, 0), items, M, n_out, K, dW=dW, db=db))
    print(f"wgrad {name:28s} {'':36s}{t*1e6:9.1f} us  {fl/t/1e12:7.1f} TFLOP/s")
    if prec == "bf16" and all(t_.size(1) % 8 == 0 for t_, _, _, _ in items):
        it16 = [(t_.to(torch.bfloat16), i_, None, 0) for t_, i_, _, _ in items]
        o16 = torch.empty(M, n_out, device=dev, dtype=torch.bfloat16)
        t = timeit(lambda: ops.linear_raw(it16, W, b, M, act, out=o16, tc=True))
        print(f"fwd16 {name:28s} {'':36s}{t*1e6:9.1f} us  {fl/t/1e12:7.1f} TFLOP/s  {2.0*M*(K+n_out)/t/1e9:7.0f} GB/s(alg)")
        dy16 = dy.to(torch.bfloat16); hm16 = hm.to(torch.bfloat16); dA16 = torch.empty(M, K, device=dev, dtype=torch.bfloat16)
        t = timeit(lambda: ops.linear_raw([(dy16, None, None, 0)], W, None, M, trans_w=True, out=dA16, out_mask=hm16, tc=True))
        print(f"dgr16 {name:28s} {'':36s}{t*1e6:9.1f} us  {fl/t/1e12:7.1f} TFLOP/s")
        t = timeit(lambda: ops.wgrad_raw((dy16, None, None, 0), it16, M, n_out, K, dW=dW, db=db, tc=True))
        print(f"wgr16 {name:28s} {'':36s}{t*1e6:9.1f} us  {fl/t/1e12:7.1f} TFLOP/s")


h1 = torch.randn(E, 256, device=dev); h2 = torch.randn(E, 128, device=dev); f1 = torch.randn(E, 192, device=dev)
a288 = torch.randn(N, 288, device=dev); h512 = torch.randn(E, 512, device=dev)
print(f"precision={prec} E={E} N={N}")
layer("edge_update.0 (gathered)", [(x, G.dst32, None, 0), (x, G.src32, None, 0), (e, None, None, 0), (att, None, None, 0)], 256)
layer("edge_update.2", [(h1, None, None, 0)], 128)
layer("edge_update.4", [(h2, None, None, 0)], 64, act=0)
layer("create_future_msgs.0 (gath)", [(x, G.dst32, None, 0), (e, None, None, 0), (x0, G.dst32, None, 0)], 192)
layer("create_future_msgs.2", [(f1, None, None, 0)], 128, act=0)
layer("att_edge_encoder.0 (gath)", [(a288[:, :64], G.dst32, None, 0), (a288[:, 64:192], G.dst32, None, 0),
                                    (a288[:, 192:], G.dst32, None, 0), (a288[:, :64], G.src32, None, 0),
                                    (a288[:, 64:192], G.src32, None, 0), (a288[:, 192:], G.src32, None, 0),
                                    (e, None, None, 0)], 512)
layer("att_edge_encoder.2", [(h512, None, None, 0)], 384)
m = torch.randn(E, 128, device=dev)
t = timeit(lambda: ops.segment_sum_raw(m, G.by_dst))
print(f"segment_sum dst [E,128]  {t*1e6:9.1f} us  {E*128*4/t/1e9:7.0f} GB/s")
t = timeit(lambda: ops.segment_sum_raw(m, G.by_src))
print(f"segment_sum src [E,128]  {t*1e6:9.1f} us  {E*128*4/t/1e9:7.0f} GB/s")
