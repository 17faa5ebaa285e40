"""Throughput of the window-file loader (GraphDataset.__getitem__ + DataLoader(batch_size=2),
batch_3dmot/utils/graph_data.py:152-256, train.py:85-97) over a generated directory of window graphs:

    python scripts/bench_loader.py [--windows 540] [--workers 8] [--dir /dev/shm/b3d_windows] [--gpu]

Writes `--windows` window file sets in the reference's on-disk layout (val-shaped: 5 frames, ~75 nodes per frame,
19-d poses, 3x32x32 image crops, 128x3 LiDAR and 64x4 radar points per node), then times
(a) the reference-style path: one window after the other through oracle/graph_io.getitem's per-edge Python loops
    (a sample of the windows), and
(b) batch3dmot_b200.graph_io.WindowBatchLoader (background threads, table look-ups instead of per-edge loops,
    collation, pinned batches), optionally (c) with the H2D copy of every batch to cuda:0 on a copy stream."""
import argparse
import os
import shutil
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from batch3dmot_b200 import graph_io, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--windows", type=int, default=540)
    ap.add_argument("--workers", type=int, default=8)
    ap.add_argument("--dir", default="/dev/shm/b3d_windows")
    ap.add_argument("--gpu", action="store_true")
    a = ap.parse_args()
    shutil.rmtree(a.dir, ignore_errors=True)
    os.makedirs(a.dir)
    prefixes, nbytes = [], 0
    t0 = time.perf_counter()
    scene_id = 0
    while len(prefixes) < a.windows:
        sc = synth.add_labels(synth.scene_graph(seed=7000 + scene_id, T=40, nodes_per_frame=75), 7000 + scene_id)
        for wi, w in enumerate(synth.windows(sc, 5)):
            if len(prefixes) >= a.windows:
                break
            n = w.pose_feats.size(0)
            w.img_feats, w.lidar_feats, w.radar_feats = torch.randn(n, 3, 32, 32), torch.randn(n, 128, 3), torch.randn(n, 64, 4)
            w.y = sc.y[w.global_edge_id]
            meta = {j: {"category_name": synth.CATEGORIES[int(w.node_classes[j]) - 1], "global_node_id": int(w.global_node_id[j])}
                    for j in range(n)}
            p = os.path.join(a.dir, f"scene{scene_id}_len5_{wi}")
            graph_io.save_window_graph(p, w, meta, boxes=torch.zeros(n, 7))
            prefixes.append(p)
        scene_id += 1
    for f in os.listdir(a.dir):
        nbytes += os.path.getsize(os.path.join(a.dir, f))
    print(f"wrote {len(prefixes)} windows, {nbytes / 1e9:.2f} GB in {time.perf_counter() - t0:.1f} s")
    from oracle import graph_io as R
    sample = prefixes[:: max(1, len(prefixes) // 20)]
    t0 = time.perf_counter()
    for p in sample:
        R.getitem(p)
    t_ref = (time.perf_counter() - t0) / len(sample)
    print(f"reference-style __getitem__ (per-edge Python loops): {1 / t_ref:.1f} windows/s ({t_ref * 1e3:.0f} ms per window)")
    for pin in (False, True):
        ld = graph_io.WindowBatchLoader(prefixes, batch_size=2, workers=a.workers, prefetch=2 * a.workers,
                                        pin=pin and torch.cuda.is_available())
        t0 = time.perf_counter()
        n_e = 0
        copy = torch.cuda.Stream() if (a.gpu and pin) else None
        for b in ld:
            n_e += b.edge_index.size(1)
            if copy is not None:
                with torch.cuda.stream(copy):
                    keep = [v.cuda(non_blocking=True) for v in vars(b).values() if torch.is_tensor(v)]
        if copy is not None:
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"WindowBatchLoader pin={pin} workers={a.workers}{' + H2D' if copy is not None else ''}: "
              f"{len(prefixes) / dt:.1f} windows/s, {n_e / dt / 1e6:.2f} M edges/s, {nbytes / dt / 1e9:.2f} GB/s of files")
    shutil.rmtree(a.dir, ignore_errors=True)


if __name__ == "__main__":
    main()
