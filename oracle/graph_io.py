"""TEST INFRASTRUCTURE ONLY (oracle): literal restatement of the reference's per-window graph loader,
GraphDataset.__getitem__ (batch_3dmot/utils/graph_data.py:152-256) with its class-balanced weight
helper cb_scaling_factor (graph_data.py:126-138), returning a plain namespace instead of a
torch_geometric.data.Data (only attribute access is used downstream). Nothing under batch3dmot_b200/
may import this module. Pinned by tests/test_graph_io.py::test_live_unmodified_reference_getitem, which runs the
unmodified reference class (imported from /root/reference under oracle/pyg_shim.py) on the same files.

Reference quirk kept: the branch for edges between nodes of DIFFERENT categories reads `self.rel_freq`,
which the reference never defines (graph_data.py:218-222 would raise AttributeError); the graph construction
only links same-category nodes, so the branch is unreachable on its own files.
"""
import json
from types import SimpleNamespace

import torch

REL_FREQ_TRAIN = {'bicycle': 0.07455396870915335, 'bus': 0.013947840246335299, 'car': 0.44736907722651076,
                  'motorcycle': 0.055813302136334404, 'pedestrian': 0.1980141158741746,
                  'trailer': 0.06407160593555014, 'truck': 0.14623008987194142}      # graph_data.py:61-68
CLASS_DICT = {'bicycle': 1, 'bus': 2, 'car': 3, 'motorcycle': 4, 'pedestrian': 5, 'trailer': 6, 'truck': 7}


def cb_scaling_factor(edge_class):                                                   # graph_data.py:126-138
    num_edges = 5
    beta = (num_edges - 1) / num_edges
    edges_per_cls = {cls: num_edges * cls_freq for cls, cls_freq in REL_FREQ_TRAIN.items()}
    return (1 - beta) / (1 - beta ** edges_per_cls[edge_class])


def getitem(prefix, inference=False, edge_weighting=True):                           # graph_data.py:152-256
    pose_features = torch.load(prefix + '_pose_features.pth')
    img_features = torch.load(prefix + '_img_features.pth')
    lidar_features = torch.load(prefix + '_lidar_features.pth')
    radar_features = torch.load(prefix + '_radar_features.pth')
    node_timestamps = torch.load(prefix + '_node_timestamps.pth')
    edge_features = torch.load(prefix + '_edge_features.pth')
    edges = torch.load(prefix + '_edges.pth')
    gt = torch.load(prefix + '_gt.pth')
    if inference:
        boxes = torch.load(prefix + '_node_boxes.pth')
    with open(prefix + '_node_metadata.json', 'r') as file:
        node_metadata = json.load(file)
    if inference:
        global_edge_index = torch.zeros_like(edges)
        global_node_timestamps = torch.zeros((node_timestamps.shape[0], 2))
        for row_idx, edge in enumerate(edges):
            global_node_j = node_metadata[str(edge[0].item())]['global_node_id']
            global_node_i = node_metadata[str(edge[1].item())]['global_node_id']
            global_edge_index[row_idx] = torch.tensor([global_node_j, global_node_i])
        for node_idx, node_time in enumerate(node_timestamps):
            global_node_timestamps[node_idx] = torch.tensor([node_metadata[str(node_idx)]['global_node_id'], node_time])
    if edge_weighting:
        weights = torch.zeros(edges.shape[0])
        edge_classes = torch.zeros(edges.shape[0])
        node_classes = torch.zeros(pose_features.shape[0])
        for row_idx, edge in enumerate(edges):
            class_a = node_metadata[str(edge[0].item())]['category_name']
            class_b = node_metadata[str(edge[1].item())]['category_name']
            if class_a == class_b:
                weights[row_idx] = cb_scaling_factor(edge_class=class_a)
                edge_classes[row_idx] = CLASS_DICT[class_a]
                node_classes[edge[0]] = CLASS_DICT[class_a]
                node_classes[edge[1]] = CLASS_DICT[class_a]
            else:
                raise AttributeError("'GraphDataset' object has no attribute 'rel_freq'")   # graph_data.py:218
    else:
        weights = torch.ones(edges.shape[0])
        edge_classes = node_classes = None
    data = SimpleNamespace(pose_feats=pose_features, img_feats=img_features, lidar_feats=lidar_features,
                           radar_feats=radar_features, edge_index=edges.t().contiguous(), edge_attr=edge_features,
                           y=gt.t().contiguous(), node_timestamps=node_timestamps, edge_weights=weights,
                           edge_classes=edge_classes, node_classes=node_classes, num_nodes=pose_features.shape[0])
    if inference:
        data.global_edge_index = global_edge_index.t().contiguous()
        data.global_node_timestamps = global_node_timestamps
        data.boxes = boxes
    return data
