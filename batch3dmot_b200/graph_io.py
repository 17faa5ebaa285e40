"""Per-window graph files -> the batch the hot path consumes: loader for the file set the
reference's graph construction writes (`<prefix>_{pose,img,lidar,radar}_features.pth`, `_node_timestamps.pth`,
`_edge_features.pth`, `_edges.pth` [E,2], `_gt.pth`, `_node_boxes.pth`, `_node_metadata.json`;
construct_detection_graph_disjoint_parallel_only_poses.py:279-308), producing what
GraphDataset.__getitem__ returns (batch_3dmot/utils/graph_data.py:152-256) without its per-edge Python
loops: class-balanced weights, edge / node classes and global ids are table look-ups over the whole edge
list. Also the writer (used for fixtures / round trips), batching (`synth.collate`) and pinned staging.
"""
import json
import random
from concurrent.futures import ThreadPoolExecutor
from types import SimpleNamespace

import torch

from .synth import CATEGORIES, cb_weights, collate

_CLASS_ID = {c: i + 1 for i, c in enumerate(CATEGORIES)}        # pose_config.yaml:122-129 (bicycle 1 .. truck 7)
_NODE_FILES = ("pose_features", "img_features", "lidar_features", "radar_features", "node_timestamps")
_EDGE_FILES = ("edge_features", "edges", "gt")


def save_window_graph(prefix, data, node_metadata, boxes=None):
    """Write one window in the reference's on-disk layout. data: namespace with pose_feats, img_feats, lidar_feats,
    radar_feats, node_timestamps, edge_attr, edge_index [2,E], y; node_metadata: {node_id: {...}} with at least
    'category_name' (and 'global_node_id' for inference files)."""
    torch.save(data.pose_feats, prefix + "_pose_features.pth")
    torch.save(data.img_feats, prefix + "_img_features.pth")
    torch.save(data.lidar_feats, prefix + "_lidar_features.pth")
    torch.save(data.radar_feats, prefix + "_radar_features.pth")
    torch.save(data.node_timestamps, prefix + "_node_timestamps.pth")
    torch.save(data.edge_attr, prefix + "_edge_features.pth")
    torch.save(data.edge_index.t().contiguous(), prefix + "_edges.pth")       # stored [E,2] = [ex_id, cur_id]
    torch.save(data.y, prefix + "_gt.pth")
    if boxes is not None:
        torch.save(boxes, prefix + "_node_boxes.pth")
    with open(prefix + "_node_metadata.json", "w") as f:
        json.dump({str(k): v for k, v in node_metadata.items()}, f)


def load_window_graph(prefix, inference=False, edge_weighting=True):
    """GraphDataset.__getitem__ (graph_data.py:152-256) for one window file set."""
    t = {k: torch.load(f"{prefix}_{k}.pth", weights_only=True) for k in _NODE_FILES + _EDGE_FILES}
    with open(prefix + "_node_metadata.json") as f:
        meta = json.load(f)
    edges = t["edges"]
    N, E = t["pose_features"].shape[0], edges.shape[0]
    src, dst = (edges[:, 0], edges[:, 1]) if E else (edges.new_zeros(0), edges.new_zeros(0))
    if edge_weighting:
        cls = torch.tensor([_CLASS_ID[meta[str(n)]["category_name"]] for n in range(N)], dtype=torch.int64)
        if E and bool((cls[src] != cls[dst]).any()):
            raise NotImplementedError("edges between different categories: the reference's own branch for them reads an "
                                      "undefined attribute (graph_data.py:218); its graph files never contain such edges")
        edge_classes = cls[src].to(torch.float32) if E else torch.zeros(0)
        weights = cb_weights(edge_classes)                       # (1-b)/(1-b**n_c) table, graph_data.py:126-138
        node_classes = torch.zeros(N)                            # only nodes that appear in an edge get a class (:211-212)
        node_classes[src] = edge_classes
        node_classes[dst] = edge_classes
    else:
        weights, edge_classes, node_classes = torch.ones(E), None, None
    data = SimpleNamespace(pose_feats=t["pose_features"], img_feats=t["img_features"], lidar_feats=t["lidar_features"],
                           radar_feats=t["radar_features"], edge_index=edges.t().contiguous(), edge_attr=t["edge_features"],
                           y=t["gt"].t().contiguous(), node_timestamps=t["node_timestamps"], edge_weights=weights,
                           edge_classes=edge_classes, node_classes=node_classes, num_nodes=N)
    if inference:
        gid = torch.tensor([meta[str(n)]["global_node_id"] for n in range(N)], dtype=edges.dtype)
        data.global_edge_index = torch.stack([gid[src], gid[dst]]) if E else edges.t().contiguous()
        data.global_node_timestamps = torch.stack([gid.to(torch.float32), t["node_timestamps"].to(torch.float32)], 1)
        data.boxes = torch.load(prefix + "_node_boxes.pth", weights_only=True)
    return data


def load_batch(prefixes, pin=True, **kw):
    """Several windows -> one disjoint batch (PyG Batch semantics, synth.collate) in pinned host memory, ready for
    `tensor.to(device, non_blocking=True)` on a copy stream (bench.py's e2e loop shows the overlap)."""
    graphs = [load_window_graph(p, **kw) for p in prefixes]
    for g in graphs:                                             # collate concatenates tensor attributes only
        for k in [k for k, v in vars(g).items() if v is None]:
            delattr(g, k)
    batch = collate(graphs)
    if pin and torch.cuda.is_available():
        for k, v in vars(batch).items():
            if torch.is_tensor(v):
                setattr(batch, k, v.pin_memory())
    return batch


class WindowBatchLoader:
    """Iterates a list of window file prefixes in batches of `batch_size` windows (the reference trains with
    torch_geometric.loader.DataLoader(dataset, batch_size=2, shuffle=True, num_workers=...), train.py:88-96):
    each batch is the disjoint union of its windows (`load_batch`), read and collated by `workers` background
    threads `prefetch` batches ahead and (on a CUDA machine) placed in pinned memory, so the consumer only
    issues `non_blocking` H2D copies. Order is deterministic for a given seed; with shuffle=False it is the list order."""

    def __init__(self, prefixes, batch_size=2, shuffle=False, seed=0, workers=4, prefetch=4, drop_last=False, pin=True,
                 **load_kw):
        self.prefixes, self.batch_size, self.shuffle, self.seed = list(prefixes), int(batch_size), shuffle, seed
        self.workers, self.prefetch, self.drop_last, self.pin, self.load_kw = workers, max(1, prefetch), drop_last, pin, load_kw
        self.epoch = 0

    def __len__(self):
        n = len(self.prefixes)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def batches(self):
        """The prefix groups of the current epoch (shuffled per epoch like a DataLoader sampler)."""
        order = list(range(len(self.prefixes)))
        if self.shuffle:
            random.Random(self.seed + self.epoch).shuffle(order)
        groups = [order[i:i + self.batch_size] for i in range(0, len(order), self.batch_size)]
        if self.drop_last and groups and len(groups[-1]) < self.batch_size:
            groups.pop()
        return [[self.prefixes[j] for j in g] for g in groups]

    def __iter__(self):
        groups = self.batches()
        self.epoch += 1
        with ThreadPoolExecutor(max_workers=max(1, self.workers)) as pool:
            pending = []
            it = iter(groups)
            for g in it:
                pending.append(pool.submit(load_batch, g, self.pin, **self.load_kw))
                if len(pending) >= self.prefetch:
                    break
            while pending:
                fut = pending.pop(0)
                nxt = next(it, None)
                if nxt is not None:
                    pending.append(pool.submit(load_batch, nxt, self.pin, **self.load_kw))
                yield fut.result()
