"""TEST INFRASTRUCTURE ONLY (oracle): literal restatement of the reference's per-window graph
construction — k-NN edge selection, edge features and ground-truth edge labels — on plain arrays
instead of nuScenes `Box` objects. Nothing under batch3dmot_b200/ may import this module.

Follows, line by line:
  batch_3dmot/utils/geo_utils.py:8-21   angle_diff
  batch_3dmot/utils/geo_utils.py:24-31  center_distance (xy only)
  batch_3dmot/utils/geo_utils.py:34-43  velocity_l2 (full 3-vector)
  batch_3dmot/utils/geo_utils.py:46-57  yaw_diff
  batch_3dmot/utils/geo_utils.py:102-115 box_volume
  batch_3dmot/utils/graph_utils.py:7-30  compute_motion_edge_feats
  batch_3dmot/utils/graph_utils.py:33-88 get_knn_nodes_in_graph
  batch_3dmot/preprocessing/construct_detection_graph_disjoint_parallel_only_poses.py:204-270
      (candidate selection per current node, k, edge emission order, GT labels, |dt| feature)

A node is a dict {'node_id', 'center' np.float64[3], 'velocity' np.float64[3], 'yaw' float (what
quaternion_yaw(box.orientation) returns), 'wlh' np.float64[3], 'category' int, 'token' int or None,
'time' int (frame index)}. Pinned: oracle/gen_golden_graph.py drives the window loop with the UNMODIFIED reference
get_knn_nodes_in_graph / compute_motion_edge_feats (imported from /root/reference) and requires bit-equal edges,
labels and float64 features before writing tests/golden/graph_build_small.pt; the loop itself (inline code of
process_chunk, not importable) is restated. Unpinned upstream: torch.topk's order of exactly tied metrics.
"""
import numpy as np
import torch


def angle_diff(x, y, period):                       # geo_utils.py:8-21
    diff = (x - y + period / 2) % period - period / 2
    if diff > np.pi:
        diff = diff - (2 * np.pi)
    return diff


def center_distance(a, b):                          # geo_utils.py:24-31
    return np.linalg.norm(a['center'][:2] - b['center'][:2])


def velocity_l2(a, b):                              # geo_utils.py:34-43
    return np.linalg.norm(a['velocity'] - b['velocity'])


def yaw_diff(a, b, period=2 * np.pi):               # geo_utils.py:46-57
    return angle_diff(a['yaw'], b['yaw'], period)


def box_volume(a):                                  # geo_utils.py:102-115
    assert all(a['wlh'] > 0)
    return np.prod(a['wlh'])


def compute_motion_edge_feats(cur_node, oth_node):  # graph_utils.py:7-30
    l2_3d_dist = center_distance(cur_node, oth_node)
    yd = np.abs(yaw_diff(cur_node, oth_node))
    vol_diff = np.log(box_volume(cur_node) / box_volume(oth_node))
    return [l2_3d_dist, yd, vol_diff]


def get_knn_nodes_in_graph(cur_node, other_nodes, k):   # graph_utils.py:33-88
    transl_3d_dists, vel_dists, yaw_dists = [], [], []
    for oth_node in other_nodes:
        l2_3d_dist = center_distance(cur_node, oth_node)
        l2_vel_dist = velocity_l2(cur_node, oth_node)
        yd = yaw_diff(cur_node, oth_node)
        transl_3d_dists.append(l2_3d_dist)
        vel_dists.append(abs(l2_vel_dist))
        yaw_dists.append(abs(yd))
    transl_3d_dists = torch.tensor(transl_3d_dists)     # np.float64 scalars -> float64 tensors
    yaw_dists = torch.tensor(yaw_dists)
    vel_dists = torch.tensor(vel_dists)
    transl_3d_dists = transl_3d_dists / torch.max(transl_3d_dists)
    yaw_dists = yaw_dists / torch.max(yaw_dists)
    vel_dists = vel_dists / torch.max(vel_dists)
    motion_dists = (1 / 2) * transl_3d_dists + (1 / 4) * yaw_dists + (1 / 4) * vel_dists
    motion_dists = motion_dists / torch.max(motion_dists)
    top_k_idcs = torch.topk(motion_dists, k, largest=False).indices.tolist()
    return [other_nodes[i] for i in top_k_idcs]


def build_window_graph(frames, top_knn=40, knn_fn=None, feat_fn=None):
    """frames: list (one entry per frame of the window, in time order) of lists of node dicts WITHOUT
    'node_id' (assigned here in emission order, construct_...:163-201). Returns edges [E,2] int64
    ([ex_id, cur_id]), gt [E] int64, edge_features [E,4] float64 — construct_...:204-270.
    knn_fn / feat_fn: replacements for get_knn_nodes_in_graph / compute_motion_edge_feats (gen_golden_graph.py
    passes the UNMODIFIED reference functions here to pin the restatements above)."""
    knn_fn = knn_fn or get_knn_nodes_in_graph
    feat_fn = feat_fn or compute_motion_edge_feats
    edges, gt_edges, edge_features = [], [], []
    past_nodes, node_id = [], 0
    for cur_nodes in frames:
        for n in cur_nodes:
            n['node_id'] = node_id
            node_id += 1
        if len(past_nodes) > 0:
            for cur in cur_nodes:
                past_categ_nodes = [p for p in past_nodes if p['category'] == cur['category']]
                k = top_knn if len(past_categ_nodes) > top_knn else len(past_categ_nodes)
                if len(past_categ_nodes) > 0:
                    knn_past_nodes = knn_fn(cur, past_categ_nodes, k)
                    for ex in knn_past_nodes:
                        edges.append([ex['node_id'], cur['node_id']])
                        if ex['token'] is not None and cur['token'] is not None:
                            if ex['token'] == cur['token']:
                                cur_ex_time_diff = abs(cur['time'] - ex['time'])
                                if cur_ex_time_diff == 1:
                                    gt_edges.append(1)
                                elif cur_ex_time_diff > 1:
                                    oth_deltas = []
                                    for oth_node in knn_past_nodes:
                                        if oth_node['time'] != ex['time'] and oth_node['token'] == cur['token']:
                                            oth_deltas.append(abs(cur['time'] - oth_node['time']))
                                    if len(oth_deltas) == 0:
                                        gt_edges.append(1)
                                    else:
                                        if np.min(oth_deltas) > cur_ex_time_diff:
                                            gt_edges.append(1)
                                        elif np.min(oth_deltas) < cur_ex_time_diff:
                                            gt_edges.append(0)
                                else:
                                    gt_edges.append(0)
                            else:
                                gt_edges.append(0)
                        else:
                            gt_edges.append(0)
                        box_feats = feat_fn(ex, cur)
                        box_feats.append(abs(cur['time'] - ex['time']))
                        edge_features.append(box_feats)
        past_nodes.extend(cur_nodes)
    return (torch.tensor(edges, dtype=torch.int64).reshape(-1, 2), torch.tensor(gt_edges, dtype=torch.int64),
            torch.tensor(edge_features, dtype=torch.float64).reshape(-1, 4))


def node_feature(ego_center, ego_wlh, ego_yaw, ego_velocity, class_id, score, val, i, num_classes=7):
    """construct_...only_poses.py:159-186 for ONE detection: returns the [1,19] row appended to pose_features."""
    ego_box_yaw = torch.from_numpy(np.array([ego_yaw]))
    feat_3d_pose = torch.cat([torch.from_numpy(ego_center).float(), torch.from_numpy(ego_wlh).float(),
                              ego_box_yaw.float(), torch.from_numpy(ego_velocity).float()], dim=0)
    feat_3d_pose = feat_3d_pose.reshape(-1, 1)
    score_feat = torch.tensor(score).reshape(-1, 1)
    class_label = torch.tensor(int(class_id))
    class_one_hot = torch.nn.functional.one_hot(class_label - 1, num_classes=num_classes)
    class_one_hot = class_one_hot.reshape(-1, 1).float()
    rel_time_tensor = torch.tensor(int(val - i)).reshape(-1, 1).float()
    return torch.cat([feat_3d_pose, class_one_hot, score_feat, rel_time_tensor], dim=0).reshape(1, -1)


# ----------------------------------------------------------------------------- synthetic windows (test inputs)
def random_window(seed, T=5, max_per_frame=40, n_cat=7, n_objects=60, p_seen=0.7, dup=False, gap_frames=()):
    """Objects random-walk over frames and are detected with probability p_seen (so instance tokens repeat with
    gaps: |dt| > 1 labels), plus false positives without a token; dup adds exact duplicates (metric ties)."""
    import math
    rng = np.random.default_rng(seed)
    cat = rng.integers(1, n_cat + 1, n_objects)
    pos = rng.uniform(-50, 50, (n_objects, 3))
    vel = rng.normal(0, 3, (n_objects, 3)); vel[:, 2] = 0
    yaw = rng.uniform(-math.pi, math.pi, n_objects)
    wlh = np.exp(rng.normal(0.5, 0.3, (n_objects, 3)))
    frames = []
    for t in range(T):
        nodes = []
        if t in gap_frames:
            frames.append(nodes); continue
        seen = np.nonzero(rng.random(n_objects) < p_seen)[0][:max_per_frame]
        for o in rng.permutation(seen):
            nodes.append({'center': pos[o] + vel[o] * 0.5 * t + rng.normal(0, 0.2, 3), 'velocity': vel[o] + rng.normal(0, 0.3, 3),
                          'yaw': float(yaw[o] + rng.normal(0, 0.05)), 'wlh': wlh[o] * np.exp(rng.normal(0, 0.02, 3)),
                          'category': int(cat[o]), 'token': int(o), 'time': 10 + t})
        for _ in range(int(rng.integers(0, 6))):            # false positives: no instance token
            nodes.append({'center': rng.uniform(-50, 50, 3), 'velocity': rng.normal(0, 3, 3), 'yaw': float(rng.uniform(-3, 3)),
                          'wlh': np.exp(rng.normal(0.5, 0.3, 3)), 'category': int(rng.integers(1, n_cat + 1)), 'token': None,
                          'time': 10 + t})
        if dup and nodes:                                  # exact duplicates: genuine metric ties
            nodes.append({k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in nodes[0].items()})
        frames.append(nodes)
    return frames


def to_tensors(frames, device="cpu"):
    """Frames of node dicts -> the tensor arguments of batch3dmot_b200.graph_build.build_window_graph."""
    nodes = [n for f in frames for n in f]
    f64 = torch.float64
    if not nodes:
        z3 = torch.zeros((0, 3), dtype=f64, device=device)
        zi = torch.zeros(0, dtype=torch.int64, device=device)
        return z3, z3, torch.zeros(0, dtype=f64, device=device), z3, zi, zi, zi
    return (torch.tensor(np.stack([n['center'] for n in nodes]), dtype=f64, device=device),
            torch.tensor(np.stack([n['velocity'] for n in nodes]), dtype=f64, device=device),
            torch.tensor([n['yaw'] for n in nodes], dtype=f64, device=device),
            torch.tensor(np.stack([n['wlh'] for n in nodes]), dtype=f64, device=device),
            torch.tensor([n['category'] for n in nodes], dtype=torch.int64, device=device),
            torch.tensor([-1 if n['token'] is None else n['token'] for n in nodes], dtype=torch.int64, device=device),
            torch.tensor([n['time'] for n in nodes], dtype=torch.int64, device=device))
