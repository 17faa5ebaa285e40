"""Window-graph construction: the step that PRODUCES the hot path's inputs —
same-category k-NN edge selection with the reference's normalised motion metric, 4-d edge features
and ground-truth edge labels — vectorised over all current nodes of a window with torch tensor ops
(runs on whatever device the inputs live on), instead of the reference's per-node Python loops.

Reference: batch_3dmot/preprocessing/construct_detection_graph_disjoint_parallel_only_poses.py:204-270
(candidates = same-category nodes of ALL earlier frames of the window, k = min(top_knn, #candidates),
edges [ex_id, cur_id] grouped by current node in id order, neighbours in torch.topk order),
batch_3dmot/utils/graph_utils.py:33-88 (metric: each of xy-distance, |yaw difference|, velocity
difference divided by its maximum over the candidates; 1/2, 1/4, 1/4 weights; divided by its maximum;
torch.topk(largest=False)), graph_utils.py:7-30 and geo_utils.py:8-57,102-115 (edge features).

Exactness: float64 throughout, same operation order as the reference. Rows whose metric holds a NaN
(0/0 when a maximum is 0) or an exact tie inside the first k+1 values — where a padded 2-D top-k
could order entries differently from the reference's 1-D call — are recomputed one by one with the
reference's own 1-D sequence of tensor ops.
"""
import math

import torch


def _angle_diff(x, y):
    """geo_utils.py:8-21 with period 2*pi (the `diff > pi` branch can never fire after the modulo)."""
    period = 2 * math.pi
    return torch.remainder(x - y + period / 2, period) - period / 2


def _norm2(d):
    """np.linalg.norm of 2-vectors (geo_utils.py:31): sqrt(x*x + y*y), elementwise so the order is fixed."""
    return (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]).sqrt()


def _norm3(d):
    """np.linalg.norm of 3-vectors (geo_utils.py:43): sequential sum of squares."""
    return ((d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]).sqrt()


def _metric_1d(center, velocity, yaw, c, cand):
    """graph_utils.py:33-88 for ONE current node c against candidate ids `cand` (1-D ops, the slow exact path)."""
    t = _norm2(center[cand, :2] - center[c, :2])
    v = _norm3(velocity[cand] - velocity[c]).abs()
    y = _angle_diff(yaw[c], yaw[cand]).abs()
    t, y, v = t / t.max(), y / y.max(), v / v.max()
    m = (1 / 2) * t + (1 / 4) * y + (1 / 4) * v
    return m / m.max()


def build_window_graph(center, velocity, yaw, wlh, category, token, frame, top_knn=40):
    """One window. Nodes are rows in emission order (frame by frame, detection order inside a frame):
    center / velocity / wlh float64 [N,3] (global frame), yaw float64 [N] (quaternion_yaw of the global box),
    category int64 [N], token int64 [N] (instance token id, -1 for None), frame int64 [N] non-decreasing.
    Returns (edges [E,2] int64 = [ex_id, cur_id], gt [E] int64, edge_features [E,4] float64)."""
    dev = center.device
    N = center.size(0)
    f64 = torch.float64
    center, velocity, yaw, wlh = center.to(f64), velocity.to(f64), yaw.to(f64), wlh.to(f64)
    empty = (torch.zeros((0, 2), dtype=torch.int64, device=dev), torch.zeros(0, dtype=torch.int64, device=dev),
             torch.zeros((0, 4), dtype=f64, device=dev))
    if N == 0:
        return empty
    ids = torch.arange(N, device=dev)
    # candidates of node c = same-category nodes of earlier frames = a PREFIX of its category's id-ordered list
    order = torch.argsort(category * N + ids)                    # stable by construction: (category, id)
    cat_sorted = category[order]
    first_of_cat = torch.searchsorted(cat_sorted, category)      # start of the node's category block
    # nodes of the same category in EARLIER frames: count entries of the block with frame < frame[c]
    key = category * (int(frame.max()) + 2) + frame              # sorted order is also sorted by (category, frame)
    cnt = torch.searchsorted(key[order], key) - first_of_cat     # [N]
    cur = torch.nonzero(cnt > 0).flatten()
    if cur.numel() == 0:
        return empty
    cmax = int(cnt[cur].max())
    k_row = cnt[cur].clamp(max=top_knn)
    kmax = int(k_row.max())
    from . import ops
    if center.is_cuda and top_knn <= 63 and ops.FEATURES["window_knn"]:
        # device path: the libb3d kernel (window_knn.cu) selects per current node without materialising the
        # [rows, candidates] matrices; exact ties are ordered by candidate position (deterministic), rows whose
        # metric holds a NaN take the reference's own 1-D sequence below
        from . import _lib as L
        ex = torch.empty((cur.numel(), kmax), dtype=torch.int64, device=dev)
        flags = torch.empty(cur.numel(), dtype=torch.int32, device=dev)
        cc, vv, yy = center.contiguous(), velocity.contiguous(), yaw.contiguous()
        first_c, cnt_c = first_of_cat[cur].contiguous(), cnt[cur].contiguous()
        L.check(L.lib().b3d_window_knn(L.ptr(cc), L.ptr(vv), L.ptr(yy), L.ptr(order), L.ptr(cur), L.ptr(first_c),
                                       L.ptr(cnt_c), cur.numel(), top_knn, kmax, L.ptr(ex), L.ptr(flags), L.stream()),
                "b3d_window_knn")
        for r in torch.nonzero(flags & 1).flatten().tolist():
            c, n_c, k = int(cur[r]), int(cnt_c[r]), int(k_row[r])
            ids_r = order[int(first_c[r]):int(first_c[r]) + n_c]
            ex[r, :k] = ids_r[torch.topk(_metric_1d(center, velocity, yaw, c, ids_r), k, largest=False).indices]
        return _finish(ex, cur, k_row, kmax, center, yaw, wlh, token, frame)
    col = torch.arange(cmax, device=dev)
    valid = col[None, :] < cnt[cur][:, None]                                     # [R, cmax]
    cand = order[(first_of_cat[cur][:, None] + col[None, :]).clamp(max=N - 1)]   # candidate node ids
    ninf, pinf = float("-inf"), float("inf")

    def rowmax(a):
        return torch.where(valid, a, torch.full_like(a, ninf)).max(1, keepdim=True).values

    t = _norm2(center[cand][..., :2] - center[cur][:, None, :2])
    v = _norm3(velocity[cand] - velocity[cur][:, None, :]).abs()
    y = _angle_diff(yaw[cur][:, None], yaw[cand]).abs()
    t, y, v = t / rowmax(t), y / rowmax(y), v / rowmax(v)
    m = (1 / 2) * t + (1 / 4) * y + (1 / 4) * v
    m = m / rowmax(m)
    has_nan = (torch.isnan(m) & valid).any(1)
    m = torch.where(valid, m, torch.full_like(m, pinf))
    kk = min(kmax + 1, cmax)
    top = torch.topk(m, kk, dim=1, largest=False)
    sel = top.indices[:, :kmax]                                                  # [R, kmax] columns
    # exact ties inside the first k+1 sorted values: the order of a padded 2-D top-k is not the reference's
    tie = torch.zeros_like(has_nan)
    if kk > 1:
        eq = top.values[:, 1:] == top.values[:, :-1]
        within = torch.arange(kk - 1, device=dev)[None, :] < k_row[:, None]
        tie = (eq & within & torch.isfinite(top.values[:, 1:])).any(1)
    for r in torch.nonzero(has_nan | tie).flatten().tolist():                    # slow exact path (rare)
        c, n_c, k = int(cur[r]), int(cnt[cur[r]]), int(k_row[r])
        m1 = _metric_1d(center, velocity, yaw, c, cand[r, :n_c])
        sel[r, :k] = torch.topk(m1, k, largest=False).indices
    ex = torch.gather(cand, 1, sel.clamp(max=cmax - 1))                          # neighbour node ids
    return _finish(ex, cur, k_row, kmax, center, yaw, wlh, token, frame)


def _finish(ex, cur, k_row, kmax, center, yaw, wlh, token, frame):
    """Selected neighbours [R, kmax] -> (edges, ground-truth labels, 4-d edge features) in emission order."""
    dev, f64 = ex.device, torch.float64
    keep = torch.arange(kmax, device=dev)[None, :] < k_row[:, None]              # [R, kmax]
    ex = torch.where(keep, ex, torch.zeros_like(ex))                             # padding -> a valid id (masked below)
    cu = cur[:, None].expand(-1, kmax)
    # ground truth (construct_...:227-260): an edge between two detections of the same instance is positive iff
    # no OTHER selected neighbour of the same instance lies in a frame closer to the current one
    same = (token[ex] == token[cu]) & (token[cu] >= 0) & (token[ex] >= 0) & keep
    dt = (frame[cu] - frame[ex]).abs()
    big = torch.iinfo(torch.int64).max
    closest = torch.where(same, dt, torch.full_like(dt, big)).min(1, keepdim=True).values
    gt = (same & (dt == closest)).to(torch.int64)
    # edge features (graph_utils.py:7-30 called as (ex, cur); construct_...:262-267 appends |dt|)
    dxy = _norm2(center[ex][..., :2] - center[cu][..., :2])
    dyaw = _angle_diff(yaw[ex], yaw[cu]).abs()
    lvol = torch.log(wlh[ex].prod(2) / wlh[cu].prod(2))
    feats = torch.stack([dxy, dyaw, lvol, dt.to(f64)], 2)
    edges = torch.stack([ex, cu], 2)
    return edges[keep], gt[keep], feats[keep]


def build_pose_features(ego_center, ego_wlh, ego_yaw, ego_velocity, category, score, rel_time, num_classes=7):
    """The 19-d node feature of construct_...only_poses.py:159-186, for all nodes at once: ego-frame center (3),
    wlh (3), yaw (1), velocity (3) as float32, one-hot class (class ids 1..num_classes), detection score, and the
    frame index relative to the window start. Inputs are [N,...] tensors (float64 box parameters as the devkit
    stores them; `.float()` rounds exactly like the reference's `torch.from_numpy(...).float()`)."""
    onehot = torch.nn.functional.one_hot(category.to(torch.int64) - 1, num_classes=num_classes).float()
    return torch.cat([ego_center.float(), ego_wlh.float(), ego_yaw.reshape(-1, 1).float(), ego_velocity.float(), onehot,
                      score.reshape(-1, 1).to(torch.float32), rel_time.reshape(-1, 1).float()], 1)
