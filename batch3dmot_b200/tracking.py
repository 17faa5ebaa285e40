"""Track assembly after the GNN: window-score averaging, per-category thresholds, greedy in/out
flux filter and the 'hier' agglomerative clustering that assigns track ids.

Vectorised re-implementation of the reference's per-scene Python dict loops
(batch_3dmot/predict.py:199-259 combine_batches_to_scene, :92-124 flux helpers, :290-373
create_trajectories mode 'hier', :437-446 track ids). The averaging / threshold / best-in /
best-out / greedy-edge steps are tensor ops on whatever device the scores live on (stable sorts
reproduce dict-insertion-order tie-breaking); only the inherently sequential clustering runs as
a compact host loop over the few surviving edges. Bit-exact against oracle/track_assembly.py."""
import numpy as np
import torch

# class ids 1..7 = bicycle, bus, car, motorcycle, pedestrian, trailer, truck (pose_config.yaml:122-129)
THRESHOLDS = torch.tensor([float("inf"), 0.1, 0.005, 0.02, 0.03, 0.025, 0.04, 0.005], dtype=torch.float64)


def average_window_scores(windows, num_nodes):
    """windows: list of (global_node_id [N_w] int64, edge_index [2,E_w] int64 (row 0 = out/source,
    row 1 = in/target; window-local), scores [E_w]). Returns (out, in, mean float64) of the
    scene-global edges in first-appearance order (== the reference's dict insertion order);
    the mean is a sequential float64 sum over the windows that contain the edge, like np.mean."""
    g_out = torch.cat([gid[ei[0]] for gid, ei, _ in windows])
    g_in = torch.cat([gid[ei[1]] for gid, ei, _ in windows])
    sc = torch.cat([s.reshape(-1) for _, _, s in windows])
    return average_edge_scores(g_out, g_in, sc, num_nodes)


def average_edge_scores(g_out, g_in, sc, num_nodes):
    """The same averaging on flat arrays: every window occurrence of an edge as (global out node, global in
    node, score), listed window after window. Node ids may span many scenes (disjoint id ranges)."""
    sc = sc.reshape(-1).to(torch.float64)
    key = g_out * num_nodes + g_in
    uniq, inv = torch.unique(key, return_inverse=True)
    occ = torch.arange(key.numel(), device=key.device)
    first = torch.full((uniq.numel(),), key.numel(), dtype=torch.long, device=key.device)
    first = first.scatter_reduce(0, inv, occ, reduce="amin")
    order = torch.argsort(first)                         # unique edges by first appearance
    rank = torch.empty_like(order)
    rank[order] = torch.arange(order.numel(), device=key.device)
    grp = rank[inv]                                      # insertion-order edge id of every occurrence
    srt = torch.sort(grp, stable=True)                   # occurrences grouped, appearance order kept
    cnt = torch.bincount(grp, minlength=order.numel())
    start = torch.cumsum(cnt, 0) - cnt
    pos = torch.arange(key.numel(), device=key.device) - start[srt.values]
    maxc = int(cnt.max()) if cnt.numel() else 0
    mat = torch.zeros((order.numel(), max(maxc, 1)), dtype=torch.float64, device=key.device)
    mat[srt.values, pos] = sc[srt.indices]
    acc = mat[:, 0].clone()
    for c in range(1, maxc):                             # sequential sum; + 0.0 is exact for short groups
        acc = acc + mat[:, c]
    mean = acc / cnt.to(torch.float64)
    u = uniq[order]
    return u // num_nodes, u % num_nodes, mean


def _first_per_group(group_key, score, n_groups):
    """For every group: index of the max-score element, ties -> smallest index (== Python's
    max(dict, key=dict.get) over insertion order). Returns [n_groups] (-1 for empty groups)."""
    by_score = torch.sort(-score, stable=True).indices
    by_group = torch.sort(group_key[by_score], stable=True)
    idx = by_score[by_group.indices]
    gk = by_group.values
    is_first = torch.ones_like(gk, dtype=torch.bool)
    is_first[1:] = gk[1:] != gk[:-1]
    out = torch.full((n_groups,), -1, dtype=torch.long, device=score.device)
    out[gk[is_first]] = idx[is_first]
    return out


def greedy_edges(e_out, e_in, mean, node_classes):
    """Threshold by the category of the OUT node (predict.py:233), keep each node's best incoming
    and best outgoing edge (:92-117) and list the survivors in the reference's greedy_edges
    insertion order (:248-255: node id order, outgoing before incoming, first insertion wins)."""
    n = node_classes.numel()
    thr = THRESHOLDS.to(mean.device)[node_classes[e_out]]
    keep = mean > thr
    e_out, e_in, mean = e_out[keep], e_in[keep], mean[keep]
    best_in = _first_per_group(e_in, mean, n)            # per node: edge id of its best predecessor edge
    best_out = _first_per_group(e_out, mean, n)
    nodes = torch.arange(n, device=mean.device)
    cand_e = torch.cat([best_out, best_in])
    cand_p = torch.cat([2 * nodes, 2 * nodes + 1])       # dict insertion position
    ok = cand_e >= 0
    cand_e, cand_p = cand_e[ok], cand_p[ok]
    pos = torch.full((mean.numel(),), 2 * n + 2, dtype=torch.long, device=mean.device)
    pos = pos.scatter_reduce(0, cand_e, cand_p, reduce="amin")
    sel = (pos < 2 * n + 2).nonzero().squeeze(1)
    sel = sel[torch.argsort(pos[sel])]
    return e_out[sel], e_in[sel], mean[sel]


def hier_tracks(e_out, e_in, score, node_classes):
    """create_trajectories(mode='hier') (predict.py:308-373) on arrays: stable descending-score order,
    then new / prepend / append / join. Returns the list of tracks in cluster insertion order."""
    j_arr = e_out.cpu().numpy(); i_arr = e_in.cpu().numpy(); s_arr = score.cpu().numpy()
    cls = node_classes.cpu().numpy()
    thr = THRESHOLDS.numpy()
    order = np.argsort(-s_arr, kind="stable")
    n = cls.shape[0]
    vis = np.full(n, -1, dtype=np.int64)
    nxt = np.full(n, -1, dtype=np.int64)
    head, tail, alive = [], [], []                       # per cluster handle, in insertion order
    for t in order:
        j, i, s = int(j_arr[t]), int(i_arr[t]), s_arr[t]
        cj, ci = vis[j], vis[i]
        if cj < 0 and ci < 0:
            head.append(j); tail.append(i); alive.append(True)
            nxt[j] = i
            vis[j] = vis[i] = len(head) - 1
        elif cj < 0:
            if head[ci] == i:
                nxt[j] = i; head[ci] = j; vis[j] = ci
        elif ci < 0:
            if tail[cj] == j:
                nxt[j] = i; tail[cj] = i; vis[i] = cj
        else:
            if tail[cj] == j and head[ci] == i and s > thr[cls[i]]:
                if cj == ci:
                    raise ValueError("edge closes a cycle inside one track (the reference code corrupts its state here)")
                nxt[j] = i
                node = i
                while node >= 0:
                    vis[node] = cj
                    node = nxt[node]
                tail[cj] = tail[ci]
                alive[ci] = False
    tracks = []
    for c in range(len(head)):
        if alive[c]:
            tr, node = [], head[c]
            while node >= 0:
                tr.append(int(node))
                node = nxt[node]
            tracks.append(tr)
    return tracks


def to_host(e_out, e_in, score, node_classes, scene_of_node=None):
    """The surviving edges and the node tables as contiguous CPU tensors (the one device->host transfer of the
    track assembly)."""
    host = [t.detach().to("cpu").contiguous() for t in
            (e_out.to(torch.int64), e_in.to(torch.int64), score.to(torch.float64), node_classes.to(torch.int64))]
    host.append(scene_of_node.detach().to("cpu", torch.int32).contiguous() if scene_of_node is not None else None)
    return host


def hier_tracks_host_arrays(e_out, e_in, score, node_classes, scene_of_node, n_scenes=1):
    """b3d_hier_tracks_host on CPU tensors (callable from a worker thread: ctypes releases the GIL)."""
    from . import _lib as L
    n = node_classes.numel()
    tid = torch.empty(n, dtype=torch.int64)
    pos = torch.empty(n, dtype=torch.int64)
    per = torch.empty(n_scenes, dtype=torch.int64)
    thr = THRESHOLDS.contiguous()
    L.check(L.lib().b3d_hier_tracks_host(L.ptr(e_out), L.ptr(e_in), L.ptr(score), e_out.numel(), L.ptr(node_classes),
                                         L.ptr(scene_of_node), n, n_scenes, L.ptr(thr), thr.numel(), L.ptr(tid), L.ptr(pos),
                                         L.ptr(per)), "b3d_hier_tracks_host")
    return tid, pos, per


def hier_tracks_native(e_out, e_in, score, node_classes, scene_of_node=None, n_scenes=1):
    """hier_tracks for the union of many scenes in ONE device->host transfer and one native loop
    (libb3d b3d_hier_tracks_host). Returns (track_id [n] int64, per-scene numbering; track_pos [n] int64;
    tracks_per_scene [n_scenes] int64) as CPU tensors."""
    return hier_tracks_host_arrays(*to_host(e_out, e_in, score, node_classes, scene_of_node), n_scenes)


def tracks_from_ids(track_id, track_pos):
    """List of tracks (each a list of node ids in order) from the arrays of hier_tracks_native (one scene)."""
    ids = track_id.numpy()
    sel = np.nonzero(ids >= 0)[0]
    order = sel[np.lexsort((track_pos.numpy()[sel], ids[sel]))]
    n_tr = int(ids.max()) + 1 if sel.size else 0
    cuts = np.searchsorted(ids[order], np.arange(1, n_tr))
    return [a.tolist() for a in np.split(order, cuts)] if n_tr else []


def assign_track_ids_union(g_out, g_in, scores, node_classes, scene_of_node=None, n_scenes=1):
    """Track assembly for the union of many scenes (node ids in disjoint ranges): flat per-window edge
    occurrences -> (track_id, track_pos, tracks_per_scene). Scenes never interact: averaging, thresholds and the
    best-in / best-out filter are per edge / per node, and the clustering visits each scene's edges in the order
    its own run would (see track_host.cu)."""
    n = node_classes.numel()
    e_out, e_in, mean = average_edge_scores(g_out, g_in, scores, n)
    k_out, k_in, k_s = greedy_edges(e_out, e_in, mean, node_classes)
    return hier_tracks_native(k_out, k_in, k_s, node_classes, scene_of_node, n_scenes)


def assign_track_ids(windows, node_classes):
    """Scene-global node id -> track id (position of its track, predict.py:438); -1 = no track."""
    n = node_classes.numel()
    e_out, e_in, mean = average_window_scores(windows, n)
    g_out, g_in, g_s = greedy_edges(e_out, e_in, mean, node_classes)
    tracks = hier_tracks(g_out, g_in, g_s, node_classes)
    ids = np.full(n, -1, dtype=np.int64)
    for tid, tr in enumerate(tracks):
        ids[tr] = tid
    return torch.from_numpy(ids), tracks
