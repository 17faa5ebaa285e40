"""Batched, scene-sharded inference (BASELINE config 4): the reference runs one GNN forward per
sliding 5-frame window and per scene inside a ray worker pool (predict.py:172-196, :595-650).
Here every rank takes whole scenes (LPT bin packing by edge count), cuts ALL windows of a chunk of its
scenes on the device with tensor ops (`window_batch`: no per-window Python loop), runs ONE forward over the
resulting disjoint batch graph (PyG-Batch style node offsets), and assembles the tracks of all scenes of the
chunk in one pass (`tracking.assign_track_ids_union`). No collective is needed: scenes are independent;
results are returned per scene id."""
from types import SimpleNamespace

import torch

from . import synth, tracking
from .parallel import lpt_partition

_NODE_KEYS = ("pose_feats", "node_timestamps", "node_classes", "x_img", "pointnet_out", "radarnet_out", "m_lidar",
              "m_radar")


def scene_windows(scene, length=5):
    """Sliding windows (stride 1) of one scene, each with `global_node_id` (predict.py:172)."""
    return synth.windows(scene, length)


def shard_scenes(scenes, world_size):
    """Scene ids per rank, balanced by total window-edge count (sizes vary ~10x between scenes)."""
    costs = [int(s.edge_index.size(1)) for s in scenes]
    return lpt_partition(costs, world_size)


_SCENE_KEYS = ("edge_attr",) + _NODE_KEYS


def pin_scene(scene):
    """Page-lock the tensors of a host-side scene (what a loader thread does once per scene), so that
    collate_scenes can issue asynchronous H2D copies instead of staged pageable ones."""
    for k, v in list(vars(scene).items()):
        if torch.is_tensor(v) and not v.is_cuda and not v.is_pinned():
            setattr(scene, k, v.pin_memory())
    return scene


def collate_scenes(scenes, device):
    """Union of whole scene graphs on `device` (node offsets in edge_index) plus what the window cutter needs:
    scene_id [N], frame index of every node inside its scene, frames per scene (host ints). Nodes of a scene
    must be stored frame after frame (true for the reference's preprocessing and for synth.scene_graph).
    Every scene tensor is copied on its own (asynchronously when it is pinned, see pin_scene) and the
    concatenation happens on the device: no host-side staging copy of the batch."""
    u = SimpleNamespace()
    on = lambda t: t.to(device, non_blocking=True)
    off, eis, sid, frame, frames = 0, [], [], [], []
    u.node_off = [0]
    for i, sc in enumerate(scenes):
        n = sc.pose_feats.size(0)
        ts = sc.node_timestamps
        if not hasattr(sc, "_t_range"):                       # host ints, computed once per scene
            sc._t_range = (int(ts.min()), int(ts.max())) if n else (0, -1)
        t0, t1 = sc._t_range
        frames.append(t1 - t0 + 1 if n else 0)
        frame.append(on(ts) - t0)
        eis.append(on(sc.edge_index) + off)
        sid.append(torch.full((n,), i, dtype=torch.long, device=device))
        off += n
        u.node_off.append(off)
    u.edge_index = torch.cat(eis, 1)
    u.scene_id = torch.cat(sid)
    u.frame = torch.cat(frame)
    for k in _SCENE_KEYS:
        if all(hasattr(sc, k) for sc in scenes):
            setattr(u, k, torch.cat([on(getattr(sc, k)) for sc in scenes]))
    u.frames, u.num_nodes, u.n_scenes = frames, off, len(scenes)
    return u


def window_batch(u, length=5):
    """Every sliding window (stride 1, `length` frames; predict.py:172) of every scene of the union `u` as ONE
    disjoint batch graph, built with tensor ops on u's device. A window's nodes are a contiguous id range
    (frames are stored one after the other), so node gathering is a range concatenation; an edge (frame fs ->
    frame fd) belongs to the windows [max(0, fd-length+1), min(fs, W-1)] of its scene and is replicated by
    repeat_interleave + one stable sort by window, which keeps the edge order of every window (targets
    ascending) and therefore of the whole batch. Equal to collate(synth.windows(scene) for every scene)."""
    dev = u.edge_index.device
    L = length
    T = torch.tensor(u.frames, dtype=torch.long)
    W = (T - L + 1).clamp(min=0)
    fbase = torch.cumsum(T, 0) - T
    wbase = torch.cumsum(W, 0) - W
    Wt, Ft = int(W.sum()), int(T.sum())
    win_scene = torch.repeat_interleave(torch.arange(len(u.frames)), W)
    win_first = (fbase[win_scene] + torch.arange(Wt) - wbase[win_scene]).to(dev)        # first frame id of each window
    W_d, fbase_d, wbase_d = W.to(dev), fbase.to(dev), wbase.to(dev)
    fid = fbase_d[u.scene_id] + u.frame
    fptr = torch.zeros(Ft + 1, dtype=torch.long, device=dev)
    fptr[1:] = torch.cumsum(torch.bincount(fid, minlength=Ft), 0)
    first_node = fptr[win_first]
    n_w = fptr[win_first + L] - first_node
    node_off = torch.cumsum(n_w, 0) - n_w
    win_of_pos = torch.repeat_interleave(torch.arange(Wt, device=dev), n_w)
    total = win_of_pos.numel()
    gather = torch.arange(total, device=dev) - node_off[win_of_pos] + first_node[win_of_pos]
    src, dst = u.edge_index[0], u.edge_index[1]
    s_e = u.scene_id[dst]
    lo = (u.frame[dst] - (L - 1)).clamp(min=0)
    hi = torch.minimum(u.frame[src], W_d[s_e] - 1)
    cnt = (hi - lo + 1).clamp(min=0)
    rep = torch.repeat_interleave(torch.arange(src.numel(), device=dev), cnt)
    start = torch.cumsum(cnt, 0) - cnt
    gwin = wbase_d[s_e[rep]] + lo[rep] + (torch.arange(rep.numel(), device=dev) - start[rep])
    order = torch.sort(gwin, stable=True)
    rep, gwin = rep[order.indices], order.values
    shift = node_off[gwin] - first_node[gwin]
    g_out, g_in = src[rep], dst[rep]
    b = SimpleNamespace(edge_index=torch.stack([g_out + shift, g_in + shift]), edge_attr=u.edge_attr[rep],
                        num_nodes=total, g_out=g_out, g_in=g_in, window_of_edge=gwin, global_node_id=gather,
                        n_windows=Wt, window_scene=win_scene, batch=win_of_pos)
    for k in _NODE_KEYS:
        if hasattr(u, k):
            setattr(b, k, getattr(u, k)[gather])
    return b


def forward_scores(model, b, multimodal=True):
    """Edge scores in (0,1) of a batch graph under torch.no_grad()."""
    with torch.no_grad():
        if multimodal:
            out, _ = model(b, x_img=b.x_img, pointnet_out=b.pointnet_out, radarnet_out=b.radarnet_out,
                           lidar_mask=b.m_lidar, radar_mask=b.m_radar)
        else:
            out, _ = model(b)
            out = torch.sigmoid(out)            # PoseGNN returns logits (pose_gnn.py:86)
    return out.reshape(-1).float()


def capture_forward(model, b, multimodal=True, warmup=2):
    """The inference forward of the batch graph `b` captured in a CUDA graph; returns `replay() -> scores [E]` (a static
    buffer, overwritten by every replay). For the launch-bound regime — one scene graph or a few window graphs per call
    (BASELINE configs[0]: a poses-only forward is ~150 short kernels for 61 k edges), where Python / ctypes dispatch
    costs several times the kernels. Shapes and addresses are static: new inputs of the SAME graph structure are copied
    into b's tensors in place (`b.pose_feats.copy_(...)`); a different edge_index needs a new capture. The weight packs
    are re-made inside the graph from the live parameters at every replay, so in-place weight updates
    (optimizer.step(), load_state_dict) are followed."""
    from . import ops
    if getattr(b, "_b3d_graph", None) is None:
        b._b3d_graph = ops.Graph(b.edge_index, b.pose_feats.size(0))
    dev = b.pose_feats.device
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for _ in range(warmup):                  # allocator pools, lazily built tables (degree blocks), module caches
            forward_scores(model, b, multimodal)
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize(dev)
    ops.invalidate_weight_cache()                # packs / stacked weights are produced inside the graph
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        scores = forward_scores(model, b, multimodal)
    ops.invalidate_weight_cache()                # nothing allocated from the graph's pool stays in a cache

    def replay():
        graph.replay()
        return scores
    replay.graph, replay.scores = graph, scores
    return replay


def infer_scene_scores(model, scenes, device, multimodal=True, window=5):
    """One forward over every window of every given scene. Returns, per scene, the list of
    (global_node_id, edge_index, scores) triples that tracking.assign_track_ids consumes
    (ids local to the scene / window, like the reference's per-window files)."""
    if not scenes:
        return []
    u = collate_scenes(scenes, device)
    b = window_batch(u, window)
    per_scene = [[] for _ in scenes]
    if b.n_windows == 0:
        return per_scene
    scores = forward_scores(model, b, multimodal) if b.edge_index.size(1) else torch.zeros(0, device=device)
    e_cnt = torch.bincount(b.window_of_edge, minlength=b.n_windows).tolist()
    n_cnt = torch.bincount(b.batch, minlength=b.n_windows).tolist()
    eo = no = 0
    for w, si in enumerate(b.window_scene.tolist()):
        e, n = e_cnt[w], n_cnt[w]
        gid = b.global_node_id[no:no + n] - u.node_off[si]
        per_scene[si].append((gid, b.edge_index[:, eo:eo + e] - no, scores[eo:eo + e]))
        eo += e
        no += n
    return per_scene


def chunk_scenes(ids, costs, max_cost):
    """Consecutive groups of scene ids whose summed cost stays below max_cost (at least one scene per group)."""
    groups, cur, tot = [], [], 0
    for i in ids:
        if cur and tot + costs[i] > max_cost:
            groups.append(cur)
            cur, tot = [], 0
        cur.append(i)
        tot += costs[i]
    if cur:
        groups.append(cur)
    return groups


def track_scenes(model, scenes, device, rank=0, world_size=1, multimodal=True, window=5, max_edges=2_500_000,
                 want_tracks=True):
    """Scene-sharded inference + track assembly. Returns {scene_id: (track_ids [N] int64, tracks)}
    for the scenes owned by `rank`. Scenes are processed in chunks of at most `max_edges` scene edges (each
    edge sits in ~2.5 windows): per chunk one window cut, one forward, one track assembly. The chunks form a
    pipeline on the host side: the copies of chunk k+1 are submitted (asynchronously, from pinned scenes) while
    chunk k is in the model, and the sequential clustering of chunk k-1 runs on a host thread
    (b3d_hier_tracks_host releases the GIL)."""
    from concurrent.futures import ThreadPoolExecutor
    costs = [int(s.edge_index.size(1)) for s in scenes]
    mine = lpt_partition(costs, world_size)[rank]
    # at least ~4 chunks per rank (when the share is big enough to split) so that the three pipeline stages overlap
    # also for small shares (many ranks, few scenes each)
    share = sum(costs[i] for i in mine)
    groups = chunk_scenes(mine, costs, min(max_edges, max(300_000, -(-share // 4))))
    out = {}

    def stage(group):
        # asynchronous copies from pinned scenes, enqueued on the compute stream behind the previous chunk's kernels:
        # what overlaps is the HOST side (this chunk's Python / copy submission under the previous chunk's GPU work)
        sub = [scenes[i] for i in group]
        return sub, collate_scenes(sub, device)

    def finish(group, sub, u, fut):
        tid, pos, _ = fut.result()
        for j, sid in enumerate(group):
            a, e = u.node_off[j], u.node_off[j + 1]
            ids = tid[a:e].clone()
            out[sid] = (ids, tracking.tracks_from_ids(ids, pos[a:e]) if want_tracks else None)

    with ThreadPoolExecutor(max_workers=1) as pool:
        staged = stage(groups[0]) if groups else None
        pending = None
        for gi, group in enumerate(groups):
            sub, u = staged
            staged = stage(groups[gi + 1]) if gi + 1 < len(groups) else None
            b = window_batch(u, window)
            if b.n_windows == 0 or b.edge_index.size(1) == 0:
                for sid, sc in zip(group, sub):
                    out[sid] = (torch.full((sc.pose_feats.size(0),), -1, dtype=torch.long), [])
                continue
            scores = forward_scores(model, b, multimodal)
            n = u.node_classes.numel()
            e_out, e_in, mean = tracking.average_edge_scores(b.g_out, b.g_in, scores, n)
            k_out, k_in, k_s = tracking.greedy_edges(e_out, e_in, mean, u.node_classes)
            host = tracking.to_host(k_out, k_in, k_s, u.node_classes, u.scene_id.int())     # one D2H of the survivors
            if pending is not None:
                finish(*pending)
            pending = (group, sub, u, pool.submit(tracking.hier_tracks_host_arrays, *host, len(sub)))
        if pending is not None:
            finish(*pending)
    return out
