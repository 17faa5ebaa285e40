"""Pins oracle/graph_construction.py against the UNMODIFIED reference utilities and writes the golden
vectors for the graph-construction row (tests/golden/graph_build_small.pt). Build container only.

    python -m oracle.gen_golden_graph

For every synthetic window the reference's own get_knn_nodes_in_graph / compute_motion_edge_feats
(batch_3dmot/utils/graph_utils.py, geo_utils.py, imported from /root/reference under oracle/pyg_shim.py's
stand-ins for pyquaternion / nuscenes Box) drive the window loop; the result must equal the restatement
exactly (edges, labels AND float64 features, bit for bit). The node yaw the restatement uses is what the
reference's quaternion_yaw returns for the box.
"""
import os

import numpy as np
import torch

from oracle import graph_construction as G
from oracle import pyg_shim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CATS = ["bicycle", "bus", "car", "motorcycle", "pedestrian", "trailer", "truck"]


def main():
    geo, gu, Box = pyg_shim.load_reference_graph_utils()
    cases = {}
    for name, kw in (("w0", dict(seed=100)), ("w1", dict(seed=101, gap_frames=(1,))), ("w2", dict(seed=102, n_cat=1, n_objects=48, p_seen=0.95, max_per_frame=45, T=4)),
                     ("w3", dict(seed=103, max_per_frame=12, n_objects=20))):
        frames = G.random_window(**kw)
        for f in frames:
            for n in f:
                n['box'] = Box(n['center'], n['wlh'], n['yaw'], n['velocity'], CATS[n['category'] - 1], n['token'])
                n['yaw'] = float(geo.quaternion_yaw(n['box'].orientation))     # what the reference's yaw_diff sees
        ref = G.build_window_graph([[dict(n) for n in f] for f in frames], knn_fn=gu.get_knn_nodes_in_graph,
                                   feat_fn=lambda ex, cur: gu.compute_motion_edge_feats(ex, cur))
        own = G.build_window_graph([[dict(n) for n in f] for f in frames])
        for a, b, what in zip(ref, own, ("edges", "gt", "edge_features")):
            assert a.dtype == b.dtype and torch.equal(a, b), f"{name}: restatement != reference ({what})"
        center, velocity, yaw, wlh, category, token, frame = G.to_tensors(frames)
        cases[name] = dict(center=center, velocity=velocity, yaw=yaw, wlh=wlh, category=category, token=token, frame=frame,
                           edges=ref[0], gt=ref[1], edge_features=ref[2])
        print(f"{name}: N = {center.size(0)}, E = {ref[0].size(0)}, positives = {int(ref[1].sum())}: reference == restatement")
    out = os.path.join(ROOT, "tests", "golden", "graph_build_small.pt")
    torch.save(cases, out)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
