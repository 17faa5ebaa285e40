"""Literal CPU restatement of the reference's track assembly: window-score
averaging, per-category thresholds, greedy in/out flux filter and the 'hier' agglomerative
clustering that yields track ids. Follows batch_3dmot/predict.py:92-124 (greedy_filter_node_flux,
aggregate_node_flux), :199-259 (combine_batches_to_scene) and :290-373 (create_trajectories,
mode 'hier'); dict insertion order and `max`/`sorted` tie-breaking are kept exactly.

Pinned by tests/test_tracking.py::test_live_unmodified_reference_track_assembly: the unmodified source text of
those predict.py functions, exec'd under stand-ins (oracle/pyg_shim.load_reference_predict_functions), returns the
same tracks on synthetic scenes, exact score ties included.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py). The node-metadata hash of the reference
(predict.py:200-208) is replaced by a provided scene-global node id per window node."""
from collections import defaultdict

import numpy as np

THRESHOLDS = {'bicycle': 0.1, 'bus': 0.005, 'car': 0.02, 'motorcycle': 0.03, 'pedestrian': 0.025,
              'trailer': 0.04, 'truck': 0.005}                       # predict.py:231 and :316


def greedy_filter_node_flux(meta):                                    # predict.py:92-117
    if len(meta['incoming']) > 1:
        pred_idx = max(meta['incoming'], key=meta['incoming'].get)
        predecessor = {pred_idx: meta['incoming'][pred_idx]}
    elif len(meta['incoming']) == 1:
        predecessor = meta['incoming']
    else:
        predecessor = {}
    if len(meta['outgoing']) > 1:
        succ_idx = max(meta['outgoing'], key=meta['outgoing'].get)
        successor = {succ_idx: meta['outgoing'][succ_idx]}
    elif len(meta['outgoing']) == 1:
        successor = meta['outgoing']
    else:
        successor = {}
    return predecessor, successor


def combine_windows(windows, node_categories):
    """windows: list of (global_node_id [N_w] ints, edges [E_w,2] (out,in) window-local, scores [E_w] float32).
    node_categories: list of category names indexed by scene-global node id (first-appearance order).
    Returns (scene_nodes_w_flux, pred_edges) like combine_batches_to_scene (predict.py:143-259)."""
    scene_edges = defaultdict(list)
    for gid, edges, scores in windows:
        for edge_idx, (out_idx, in_idx) in enumerate(edges):
            scene_edges[(int(gid[out_idx]), int(gid[in_idx]))].append(float(scores[edge_idx]))      # :221
    scene_nodes = {i: {'category_name': c, 'incoming': dict(), 'outgoing': dict()} for i, c in enumerate(node_categories)}
    avg_edge_scores = {edge: np.mean(scores) for edge, scores in scene_edges.items()}               # :227
    avg_edge_scores = {edge: s for edge, s in avg_edge_scores.items()
                       if s > THRESHOLDS[scene_nodes[edge[0]]['category_name']]}                    # :233
    for (out_idx, in_idx), score in avg_edge_scores.items():                                        # :120-124
        scene_nodes[in_idx]['incoming'].update({out_idx: float(score)})
        scene_nodes[out_idx]['outgoing'].update({in_idx: float(score)})
    for node_idx, node_meta in scene_nodes.items():                                                 # :243-245
        scene_nodes[node_idx]['incoming'], scene_nodes[node_idx]['outgoing'] = greedy_filter_node_flux(node_meta)
    greedy_edges = dict()
    for node_idx in scene_nodes:                                                                    # :248-255
        if len(scene_nodes[node_idx]['outgoing']) > 0:
            greedy_edges[(node_idx, list(scene_nodes[node_idx]['outgoing'].keys())[0])] = \
                list(scene_nodes[node_idx]['outgoing'].values())[0]
        if len(scene_nodes[node_idx]['incoming']) > 0:
            greedy_edges[(list(scene_nodes[node_idx]['incoming'].keys())[0], node_idx)] = \
                list(scene_nodes[node_idx]['incoming'].values())[0]
    pred_edges = [(edge, score) for edge, score in greedy_edges.items()]
    return scene_nodes, pred_edges


def create_trajectories(pred_edges, scene_nodes):                      # predict.py:290-373, mode 'hier'
    pred_edges_dict = {e[0]: e[1] for e in pred_edges}
    pred_edges_desc = {k: v for k, v in sorted(pred_edges_dict.items(), key=lambda item: item[1], reverse=True)}
    vis = defaultdict()
    clusters = defaultdict(list)
    clusters_scores = defaultdict(list)
    for edge, score in pred_edges_desc.items():
        j, i = edge
        edge_cat = scene_nodes[i]['category_name']
        if j not in vis.keys() and i not in vis.keys():
            cluster_idx = 0 if len(list(clusters.keys())) == 0 else max(list(clusters.keys())) + 1
            clusters[cluster_idx].extend([j, i])
            clusters_scores[cluster_idx].append(score)
            vis[i] = cluster_idx
            vis[j] = cluster_idx
        else:
            if j not in vis.keys() and i in vis.keys():
                c = vis[i]
                if clusters[c][0] == i:
                    clusters[c].insert(0, j)
                    clusters_scores[c].insert(0, score)
                    vis[j] = c
                else:
                    continue
            elif j in vis.keys() and i not in vis.keys():
                c = vis[j]
                if clusters[c][-1] == j:
                    clusters[c].append(i)
                    clusters_scores[c].append(score)
                    vis[i] = c
                else:
                    continue
            elif j in vis.keys() and i in vis.keys():
                c0, c1 = vis[j], vis[i]
                if j == clusters[c0][-1] and i == clusters[c1][0] and score > THRESHOLDS[edge_cat]:
                    clusters[c0] = clusters[c0] + clusters[c1]
                    clusters_scores[c0] = clusters_scores[c0] + clusters_scores[c1]
                    for node in clusters[c0]:
                        vis[node] = c0
                    del clusters[c1]
                    del clusters_scores[c1]
                else:
                    continue
    return [v for k, v in clusters.items()]


def track_ids(windows, node_categories):
    """Scene-global node id -> track id (= position of its track in the list, predict.py:438);
    -1 for nodes that no kept edge touches."""
    nodes, pred_edges = combine_windows(windows, node_categories)
    tracks = create_trajectories(pred_edges, nodes)
    out = np.full(len(node_categories), -1, dtype=np.int64)
    for tid, tr in enumerate(tracks):
        for n in tr:
            out[n] = tid
    return out, tracks
