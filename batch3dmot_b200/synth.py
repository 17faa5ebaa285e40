"""Deterministic synthetic nuScenes-shaped tracking graphs.

The shapes follow the reference's graph construction
(batch_3dmot/preprocessing/construct_detection_graph_disjoint_parallel_only_poses.py:148-290):
19-d pose features, edges from every node to its <= k nearest same-category
nodes in the previous `window` frames, emitted grouped by target node in
ascending id (so edge_index[1] is non-decreasing and src < dst), 4-d edge
features [xy-dist, |dyaw|, log(vol_j/vol_i), dt] as float64.

All RNG is `torch.Generator().manual_seed(seed)` on CPU.
"""
import math
from types import SimpleNamespace

import torch

# class ids 1..7 in pose_config.yaml:122-129 order; frequencies graph_data.py:61-68
CATEGORIES = ["bicycle", "bus", "car", "motorcycle", "pedestrian", "trailer", "truck"]
REL_FREQ_TRAIN = {
    "bicycle": 0.07455396870915335, "bus": 0.013947840246335299, "car": 0.44736907722651076,
    "motorcycle": 0.055813302136334404, "pedestrian": 0.1980141158741746,
    "trailer": 0.06407160593555014, "truck": 0.14623008987194142,
}
_WLH_MEAN = {"bicycle": (0.6, 1.7, 1.3), "bus": (2.9, 11.0, 3.5), "car": (1.9, 4.6, 1.7),
             "motorcycle": (0.8, 2.1, 1.5), "pedestrian": (0.7, 0.7, 1.8),
             "trailer": (2.9, 12.0, 3.9), "truck": (2.5, 6.9, 2.8)}


def _wrap(a):
    return (a + math.pi) % (2 * math.pi) - math.pi


def scene_graph(seed=5621, T=40, nodes_per_frame=50, window=4, k=40, rel_time_mod=None,
                frame_sizes=None):
    """One scene (or window) graph. Returns a SimpleNamespace with the PyG `Data`
    attribute names the reference reads (pose_gnn.py:59-65): pose_feats [N,19] f32,
    edge_index [2,E] i64, edge_attr [E,4] f64, node_timestamps [N] i64, batch [N] i64,
    plus node_classes [N] i64 (1..7) and num_nodes."""
    g = torch.Generator().manual_seed(seed)
    probs = torch.tensor([REL_FREQ_TRAIN[c] for c in CATEGORIES], dtype=torch.float64)
    if frame_sizes is None:
        frame_sizes = [nodes_per_frame] * T
    feats, cats, ts, xy_all, yaw_all, vol_all = [], [], [], [], [], []
    for t, n_t in enumerate(frame_sizes):
        cat = torch.multinomial(probs, n_t, replacement=True, generator=g)
        # ego radius in (1, 50) m  (construct_..._only_poses.py:148-149)
        r = torch.sqrt(torch.rand(n_t, generator=g, dtype=torch.float64) * (50.0 ** 2 - 1.0) + 1.0)
        phi = torch.rand(n_t, generator=g, dtype=torch.float64) * 2 * math.pi
        xy = torch.stack([r * torch.cos(phi), r * torch.sin(phi)], 1)
        z = torch.randn(n_t, 1, generator=g, dtype=torch.float64)
        mean = torch.tensor([_WLH_MEAN[CATEGORIES[c]] for c in cat.tolist()], dtype=torch.float64)
        wlh = mean * torch.exp(0.1 * torch.randn(n_t, 3, generator=g, dtype=torch.float64))
        yaw = (torch.rand(n_t, generator=g, dtype=torch.float64) * 2 - 1) * math.pi
        vel = torch.cat([3.0 * torch.randn(n_t, 2, generator=g, dtype=torch.float64),
                         torch.zeros(n_t, 1, dtype=torch.float64)], 1)
        u = torch.rand(n_t, 2, generator=g, dtype=torch.float64)
        score = 0.5 * (u[:, 0] + u[:, 1])  # symmetric on (0,1), mean .5
        onehot = torch.nn.functional.one_hot(cat, 7).to(torch.float64)
        rel_t = float(t % rel_time_mod) if rel_time_mod else float(t)
        f = torch.cat([xy, z, wlh, yaw[:, None], vel, onehot, score[:, None],
                       torch.full((n_t, 1), rel_t, dtype=torch.float64)], 1)
        feats.append(f); cats.append(cat); ts.append(torch.full((n_t,), t, dtype=torch.long))
        xy_all.append(xy); yaw_all.append(yaw); vol_all.append(wlh.prod(1))
    pose = torch.cat(feats); cat = torch.cat(cats); ts = torch.cat(ts)
    xy = torch.cat(xy_all); yaw = torch.cat(yaw_all); vol = torch.cat(vol_all)
    starts = [0]
    for n_t in frame_sizes:
        starts.append(starts[-1] + n_t)

    src_l, dst_l = [], []
    for t in range(1, len(frame_sizes)):
        lo, hi = starts[max(0, t - window)], starts[t]
        c0, c1 = starts[t], starts[t + 1]
        if c1 == c0 or hi == lo:
            continue
        d = torch.cdist(xy[c0:c1], xy[lo:hi])
        d = torch.where(cat[c0:c1, None] == cat[None, lo:hi], d, torch.full_like(d, float("inf")))
        order = torch.sort(d, dim=1, stable=True)
        cnt = torch.isfinite(d).sum(1).clamp(max=k)
        kk = min(k, hi - lo)
        keep = torch.arange(kk)[None, :] < cnt[:, None]
        src = (order.indices[:, :kk] + lo)[keep]
        dst = (torch.arange(c0, c1)[:, None].expand(-1, kk))[keep]
        src_l.append(src); dst_l.append(dst)
    if src_l:
        src = torch.cat(src_l); dst = torch.cat(dst_l)
    else:
        src = dst = torch.zeros(0, dtype=torch.long)
    edge_index = torch.stack([src, dst]).contiguous()
    dxy = (xy[src] - xy[dst]).norm(dim=1)
    dyaw = _wrap(yaw[src] - yaw[dst]).abs()
    lvol = torch.log(vol[src] / vol[dst])
    dt = (ts[dst] - ts[src]).abs().to(torch.float64)
    edge_attr = torch.stack([dxy, dyaw, lvol, dt], 1)
    N = pose.size(0)
    return SimpleNamespace(
        pose_feats=pose.to(torch.float32), edge_index=edge_index, edge_attr=edge_attr,
        node_timestamps=ts, batch=torch.zeros(N, dtype=torch.long),
        node_classes=cat + 1, num_nodes=N)


def add_modalities(data, seed=5621, p_lidar=0.7, p_radar=0.3, raw=True):
    """B.2: random encoder outputs and Bernoulli modality-dropout masks. Camera is never
    masked (the reference never masks it, clr_att_gnn.py:125). Adds x_img [N,96] (ResNetAE
    .encode output), pointnet_out [N,256], radarnet_out [N,256] (encoder outputs; rows of
    missing modalities are never read), m_lidar / m_radar bool [N]. With raw=True also adds
    raw `lidar_feats [N,128,3]` / `radar_feats [N,64,4]` that are all-zero exactly where the
    modality is missing (mask derivation clr_att_gnn.py:111-121) and carry node id + 1 in
    element [n,0,0] so `EmbeddingEncoder` stubs can look rows up, and `img_feats` = node id."""
    g = torch.Generator().manual_seed(seed + 1)
    N = data.pose_feats.size(0)
    data.m_lidar = torch.rand(N, generator=g) < p_lidar
    data.m_radar = torch.rand(N, generator=g) < p_radar
    data.x_img = torch.randn(N, 96, generator=g)
    data.pointnet_out = torch.randn(N, 256, generator=g)
    data.radarnet_out = torch.randn(N, 256, generator=g)
    if raw:
        add_raw_feats(data)
    return data


def add_raw_feats(data):
    """Raw-tensor view of the modality masks (see add_modalities)."""
    N = data.pose_feats.size(0)
    dev = data.pose_feats.device
    ids = torch.arange(1, N + 1, dtype=torch.float32, device=dev)
    lf = torch.zeros(N, 128, 3, device=dev); lf[:, 0, 0] = ids * data.m_lidar
    rf = torch.zeros(N, 64, 4, device=dev); rf[:, 0, 0] = ids * data.m_radar
    data.lidar_feats, data.radar_feats = lf, rf
    data.img_feats = torch.arange(N, device=dev)
    return data


class EmbeddingEncoder(torch.nn.Module):
    """Stub for the frozen, out-of-scope modality encoders (resnet_fully_conv.ResNetAE,
    pointnet.PointNetClassifier, radarnet.RadarNetClassifier): returns pre-drawn rows.
    Duck-types what GNN touches: .encode / .forward_feat / .named_parameters / .eval
    (clr_att_gnn.py:26-33,125-141)."""

    def __init__(self, table):
        super().__init__()
        self.register_buffer("table", table, persistent=False)

    def encode(self, img_feats):            # ResNetAE.encode: img_feats holds node ids
        return self.table[img_feats.view(-1).long()]

    def forward_feat(self, feats):          # feats [n,C,P] view of rows carrying id+1 at [n,0,0]
        return self.table[feats[:, 0, 0].long() - 1]


def add_labels(data, seed=5621, p_pos=0.03):
    """Config 5: y ~ Bernoulli(p_pos); class-balanced weights per A.7 (graph_data.py:126-138)."""
    g = torch.Generator().manual_seed(seed + 2)
    E = data.edge_index.size(1)
    data.y = (torch.rand(E, generator=g) < p_pos).to(torch.long)
    data.edge_classes = data.node_classes[data.edge_index[1]].to(torch.float32)
    data.edge_weights = cb_weights(data.edge_classes)
    return data


def cb_weights(edge_classes):
    """Vectorised class-balanced weights: w = (1-b)/(1-b**n_c), b = 0.8, n_c = 5*rel_freq
    (graph_data.py:126-138); replaces the per-edge Python loop at graph_data.py:202-226."""
    n_edges = 5
    beta = (n_edges - 1) / n_edges
    table = torch.tensor([0.0] + [(1 - beta) / (1 - beta ** (n_edges * REL_FREQ_TRAIN[c]))
                                  for c in CATEGORIES], dtype=torch.float64)
    return table[edge_classes.to(torch.long)].to(torch.float32)


def collate(graphs):
    """PyG `Batch.from_data_list` semantics for the attributes the hot path reads: node
    tensors concatenated along dim 0, edge_index along dim 1 with node offsets, `batch`
    = graph id per node (train.py:88-96 via torch_geometric.loader.DataLoader)."""
    out = SimpleNamespace()
    off, eis, batch = 0, [], []
    for gi, gph in enumerate(graphs):
        n = gph.pose_feats.size(0)
        eis.append(gph.edge_index + off)
        batch.append(torch.full((n,), gi, dtype=torch.long))
        off += n
    out.edge_index = torch.cat(eis, 1).contiguous()
    out.batch = torch.cat(batch)
    out.num_nodes = off
    node_edge_keys = [k for k, v in vars(graphs[0]).items()
                      if torch.is_tensor(v) and k not in ("edge_index", "batch")]
    for k in node_edge_keys:
        # [2,E]-shaped index pairs (`global_edge_index` of the inference loader, graph_data.py:244-250) are
        # concatenated along the edge dimension like edge_index. They hold SCENE-global node ids, so they are
        # NOT offset (PyG's Batch would add the window's node offset to any key containing "index", which
        # would corrupt the global ids; predict.py never batches windows, so the reference has no behaviour here).
        dim = 1 if (k.endswith("edge_index") and getattr(graphs[0], k).dim() == 2
                    and getattr(graphs[0], k).size(0) == 2) else 0
        setattr(out, k, torch.cat([getattr(gph, k) for gph in graphs], dim))
    return out


def windows(scene, length=5):
    """Cut a scene graph into sliding windows of `length` frames, stride 1
    (predict.py:172), re-indexing nodes; node_timestamps stay absolute. Each window
    carries `global_node_id` (scene-level node id) for track assembly (A.8)."""
    ts = scene.node_timestamps
    T = int(ts.max()) + 1 if ts.numel() else 0
    out = []
    src, dst = scene.edge_index
    for t0 in range(0, T - length + 1):
        nm = (ts >= t0) & (ts < t0 + length)
        ids = nm.nonzero().squeeze(1)
        remap = torch.full((ts.numel(),), -1, dtype=torch.long)
        remap[ids] = torch.arange(ids.numel())
        em = nm[src] & nm[dst]
        w = SimpleNamespace(
            pose_feats=scene.pose_feats[ids], edge_index=torch.stack([remap[src[em]], remap[dst[em]]]),
            edge_attr=scene.edge_attr[em], node_timestamps=ts[ids],
            batch=torch.zeros(ids.numel(), dtype=torch.long), node_classes=scene.node_classes[ids],
            num_nodes=ids.numel(), global_node_id=ids, global_edge_id=em.nonzero().squeeze(1))
        for k in ("x_img", "pointnet_out", "radarnet_out", "m_lidar", "m_radar"):
            if hasattr(scene, k):
                setattr(w, k, getattr(scene, k)[ids])
        out.append(w)
    return out


def knn_stress(seed, N, frame=250, D=48, skewed=False, continuous=False):
    """B.3: features on the 2^-5 grid in [-4,4) (all squared distances exact in fp32) or
    N(0,1); frames of `frame` nodes or LogNormal-skewed sizes in [1, 5000]. Returns
    (x [N,D] f32, frame_ptr [F+1] i64); nodes are grouped by frame."""
    g = torch.Generator().manual_seed(seed)
    if continuous:
        x = torch.randn(N, D, generator=g)
    else:
        x = torch.randint(-128, 128, (N, D), generator=g).to(torch.float32) / 32.0
    if skewed:
        sizes, tot = [], 0
        while tot < N:
            s = int(torch.exp(torch.randn(1, generator=g) * 1.2 + math.log(frame)).clamp(1, 5000))
            s = min(s, N - tot); sizes.append(s); tot += s
    else:
        sizes = [frame] * (N // frame) + ([N % frame] if N % frame else [])
    ptr = torch.tensor([0] + sizes, dtype=torch.long).cumsum(0)
    return x, ptr
