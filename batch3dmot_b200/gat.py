"""Frame-wise k-NN + attention-weighted convolution (reference: torch_geometric.nn.knn_graph +
GATConv at pose_gnn.py:55,75-80 and clr_att_gnn.py:93,179-184; semantics SURVEY A.5/A.6)."""
import math

import torch
from torch import nn

from . import ops


def frame_ptr_from_timestamps(node_timestamps):
    """int32 [F+1] frame offsets of nodes STORED GROUPED by timestamp (runs of equal values). The reference's
    grouping key is the timestamp value alone (pose_gnn.py:76-77: `for t in node_timestamps.unique()`), so
    callers with arbitrary node order go through `knn_attention_conv`, which sorts first."""
    _, counts = torch.unique_consecutive(node_timestamps, return_counts=True)
    ptr = torch.zeros(counts.numel() + 1, dtype=torch.int32, device=node_timestamps.device)
    ptr[1:] = counts.cumsum(0)
    return ptr


class GATConv(nn.Module):
    """Parameter container + forward for GATConv(D, D, heads=1, add_self_loops=False) over a
    padded neighbour table. state_dict keys follow PyG 2.0.x: lin_src.weight (shared with
    lin_dst), att_src, att_dst, bias; `lin_dst.weight` / `lin.weight` spellings are accepted
    on load."""

    def __init__(self, in_channels, out_channels, add_self_loops=False, negative_slope=0.2):
        super().__init__()
        assert not add_self_loops, "the reference only uses add_self_loops=False"
        self.in_channels, self.out_channels, self.negative_slope = in_channels, out_channels, negative_slope
        self.lin_src = nn.Linear(in_channels, out_channels, bias=False)
        self.att_src = nn.Parameter(torch.empty(1, 1, out_channels))
        self.att_dst = nn.Parameter(torch.empty(1, 1, out_channels))
        self.bias = nn.Parameter(torch.zeros(out_channels))
        for t in (self.lin_src.weight, self.att_src, self.att_dst):   # glorot
            a = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
            with torch.no_grad():
                t.uniform_(-a, a)

    def _save_to_state_dict(self, destination, prefix, keep_vars):
        """PyG 2.0.x registers the shared Linear twice (`self.lin_dst = self.lin_src`), so its state_dict
        holds `lin_src.weight` AND `lin_dst.weight`; emit both so a checkpoint written here loads strictly in
        the reference (predict.py:404)."""
        super()._save_to_state_dict(destination, prefix, keep_vars)
        w = self.lin_src.weight
        destination[prefix + "lin_dst.weight"] = w if keep_vars else w.detach()

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        for alt in ("lin_dst.weight", "lin.weight"):
            k = prefix + alt
            if k in state_dict:
                v = state_dict.pop(k)
                state_dict.setdefault(prefix + "lin_src.weight", v)
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)

    def forward_table(self, x, nbr):
        """x [N,D], nbr int64 [N,k] (-1 padded; row t lists the sources of target t)."""
        h = ops.fused_linear([(x, None)], self.lin_src.weight)
        return ops.gat_aggregate(h, self.att_src, self.att_dst, self.bias, nbr, self.negative_slope)

    def forward(self, x, edge_index):
        """PyG signature for knn_graph edge lists (row 0 source, row 1 target, grouped by target)."""
        N = x.size(0)
        t = edge_index[1]
        deg = torch.bincount(t, minlength=N)
        k = max(int(deg.max()) if deg.numel() else 0, 1)
        start = torch.cumsum(deg, 0) - deg
        pos = torch.arange(t.numel(), device=t.device) - start[t]
        nbr = torch.full((N, k), -1, dtype=torch.int64, device=x.device)
        nbr[t, pos] = edge_index[0]
        return self.forward_table(x, nbr)


def knn_attention_conv(conv, x, node_timestamps, k=20, frame_ptr=None):
    """The frame-wise k-NN graph + GATConv pass (pose_gnn.py:76-79): returns the updated
    node features. The reference DISCARDS this result (`==` at pose_gnn.py:80, quirk C1).

    Frames are the sets of nodes with EQUAL timestamp wherever they are stored (the reference masks
    `node_timestamps == t`): nodes are stably sorted by timestamp, the kernel runs on the grouped copy
    and the neighbour table is mapped back, so ungrouped inputs give the reference's frames and the
    (distance, node id) tie order is unchanged (a stable sort keeps ids ascending inside a frame).
    Pass `frame_ptr` to assert that nodes are already grouped and skip the sort."""
    if frame_ptr is not None:
        return conv.forward_table(x, ops.knn_frames(x, frame_ptr, k))
    ts = node_timestamps.reshape(-1)
    order = torch.sort(ts, stable=True).indices
    frame_ptr = frame_ptr_from_timestamps(ts[order])
    nbr_s = ops.knn_frames(x.detach()[order], frame_ptr, k)            # ids in the sorted numbering
    nbr = torch.where(nbr_s >= 0, order[nbr_s.clamp(min=0)], nbr_s)   # -> original node ids
    table = torch.empty_like(nbr)
    table[order] = nbr                                                # row of sorted node p belongs to node order[p]
    return conv.forward_table(x, table)
