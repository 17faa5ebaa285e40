"""Window-graph construction: the reference's per-node loops (literal restatement, CPU) against
batch3dmot_b200.graph_build on CPU tensors and, when a GPU is present, on CUDA tensors. Same synthetic windows as
tests/test_graph_build.py; results are checked for equality before timing."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import graph_construction as G
from batch3dmot_b200 import graph_build
from tests.test_graph_build import random_window, to_tensors

for per_frame, n_obj in ((75, 110), (300, 430)):
    frames = random_window(1, max_per_frame=per_frame, n_objects=n_obj, p_seen=0.7)
    n = sum(len(f) for f in frames)
    t0 = time.perf_counter(); e_ref, gt_ref, f_ref = G.build_window_graph(frames); t_ref = time.perf_counter() - t0
    line = f"window of 5 frames, N = {n} nodes, E = {e_ref.size(0)} edges: reference loops {t_ref * 1e3:.0f} ms"
    for dev in ["cpu"] + (["cuda"] if torch.cuda.is_available() else []):
        args = to_tensors(frames, dev)
        e, gt, f = graph_build.build_window_graph(*args)
        assert torch.equal(e.cpu(), e_ref) and torch.equal(gt.cpu(), gt_ref)
        reps = 20
        if dev == "cuda": torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps): graph_build.build_window_graph(*args)
        if dev == "cuda": torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        line += f"; vectorised on {dev} {dt * 1e3:.2f} ms ({t_ref / dt:.0f}x)"
    print(line)
