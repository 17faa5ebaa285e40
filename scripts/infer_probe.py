"""Stage timing of the batched inference + track assembly path (configs[3])."""
import math
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from batch3dmot_b200 import inference, ops, synth, tracking  # noqa: E402
from batch3dmot_b200.clr_att_gnn import GNN  # noqa: E402
from batch3dmot_b200.parallel import lpt_partition  # noqa: E402

dev = torch.device("cuda")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 150
g = torch.Generator().manual_seed(5621 + 77)
rates = torch.exp(torch.randn(n, generator=g) * 0.5 + math.log(75.0) - 0.125).clamp(5, 300).round().long().tolist()
t0 = time.perf_counter()
scenes = [synth.add_modalities(synth.scene_graph(seed=10621 + i, T=40, frame_sizes=[rates[i]] * 40), 10621 + i, raw=False)
          for i in range(n)]
print(f"generated {n} scenes in {time.perf_counter() - t0:.1f} s")
ops.set_precision("bf16")
torch.manual_seed(5621)
model = GNN(None, None, None).to(dev).eval()
costs = [int(s.edge_index.size(1)) for s in scenes]


def sync():
    torch.cuda.synchronize()
    return time.perf_counter()


for rep in range(2):
    T = dict(collate=0.0, windows=0.0, forward=0.0, assemble_gpu=0.0, cluster_host=0.0)
    t_all = sync()
    for group in inference.chunk_scenes(list(range(n)), costs, 2_500_000):
        sub = [scenes[i] for i in group]
        t = sync(); u = inference.collate_scenes(sub, dev); T["collate"] += sync() - t
        t = sync(); b = inference.window_batch(u, 5); T["windows"] += sync() - t
        t = sync(); scores = inference.forward_scores(model, b, True); T["forward"] += sync() - t
        t = sync()
        e_out, e_in, mean = tracking.average_edge_scores(b.g_out, b.g_in, scores, u.node_classes.numel())
        k_out, k_in, k_s = tracking.greedy_edges(e_out, e_in, mean, u.node_classes)
        T["assemble_gpu"] += sync() - t
        t = sync(); tracking.hier_tracks_native(k_out, k_in, k_s, u.node_classes, u.scene_id.int(), len(sub)); T["cluster_host"] += sync() - t
    tot = sync() - t_all
    print(f"rep {rep}: total {tot:.3f} s  " + "  ".join(f"{k} {v:.3f}" for k, v in T.items()), f" window-edges {b.edge_index.size(1)} in last chunk")

# the pipelined public path (pinned scenes, copy stream, host clustering on a worker thread)
for sc in scenes:
    inference.pin_scene(sc)
for rep in range(6):
    t = sync()
    res = inference.track_scenes(model, scenes, dev, want_tracks=False)
    dt = sync() - t
    print(f"track_scenes (pipelined, pinned): {dt:.3f} s = {n / dt:.1f} scenes/s")
