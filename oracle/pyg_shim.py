"""Minimal stand-in for the un-vendored third-party graph packages so that the
UNMODIFIED reference model files import and run on CPU.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates, from their published semantics (versions unpinned by the reference:
no requirements/lock file; PyG 2.0.x-2.1.x inferred from the private
`__check_input__/__collect__` API used at pose_gnn.py:163-181):

  torch_geometric.nn.MessagePassing  -> exactly what `propagate` touches
                                        (pose_gnn.py:134-196, clr_att_gnn.py:238-300)
  torch_geometric.nn.GATConv         -> SURVEY.md §A.6  (pose_gnn.py:55, clr_att_gnn.py:93)
  torch_geometric.nn.knn_graph       -> SURVEY.md §A.5  (pose_gnn.py:78, clr_att_gnn.py:182)
  torch_scatter.scatter              -> SURVEY.md §A.4  (pose_gnn.py:240, clr_att_gnn.py:344)

`install()` pre-populates sys.modules; `load_reference()` then imports the
reference modules from /root/reference (only possible in the build container).
"""
import inspect
import math
import sys
import types

import torch
from torch import nn


# --------------------------------------------------------------------------- scatter
def scatter(src, index, dim=0, out=None, dim_size=None, reduce="sum"):
    """torch_scatter.scatter for reduce in {'add','sum'} along dim 0 (A.4):
    out = zeros(dim_size, C); out.scatter_add_(0, index[:,None].expand_as(src), src).
    CPU scatter_add_ is sequential in edge order == index_add_."""
    assert reduce in ("add", "sum") and dim in (0, -2)
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    res = src.new_zeros((dim_size,) + tuple(src.shape[1:]))
    return res.index_add_(0, index, src)


# --------------------------------------------------------------------------- knn_graph
def knn_graph(x, k, batch=None, loop=False, flow="source_to_target", **_):
    """A.5 with the build's tie-break spec: squared L2 accumulated in fp32 in
    ascending feature order, order by (distance, neighbour index), self
    excluded by index, k_eff = min(k, n-1). Returns [2, n*k_eff] int64 with
    row 0 = neighbour (source), row 1 = query (target), grouped by query."""
    assert batch is None and not loop and flow == "source_to_target"
    n = x.size(0)
    if n <= 1:
        return torch.zeros((2, 0), dtype=torch.long, device=x.device)
    xd = x.detach().to(torch.float32)
    d = torch.zeros((n, n), dtype=torch.float32)
    for c in range(xd.size(1)):  # ascending d, separate mul and add (no FMA)
        diff = xd[:, c].unsqueeze(1) - xd[:, c].unsqueeze(0)
        d = d + diff * diff
    d.fill_diagonal_(float("inf"))
    keff = min(k, n - 1)
    # stable sort on distance == (distance, index) order
    order = torch.sort(d, dim=1, stable=True).indices[:, :keff]
    row = order.reshape(-1)
    col = torch.arange(n).unsqueeze(1).expand(n, keff).reshape(-1)
    return torch.stack([row, col], dim=0)


# --------------------------------------------------------------------------- GATConv
def _glorot(t):
    a = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
    with torch.no_grad():
        t.uniform_(-a, a)


class GATConv(nn.Module):
    """A.6: heads=1, concat=True, negative_slope=0.2, dropout=0, bias=True.
    Parameter names follow PyG 2.0.x (`lin_src`, shared with lin_dst)."""

    def __init__(self, in_channels, out_channels, heads=1, concat=True, negative_slope=0.2,
                 dropout=0.0, add_self_loops=True, bias=True, **_):
        super().__init__()
        assert heads == 1 and not add_self_loops
        self.in_channels, self.out_channels = in_channels, out_channels
        self.negative_slope = negative_slope
        self.lin_src = nn.Linear(in_channels, out_channels, bias=False)
        self.lin_dst = self.lin_src
        self.att_src = nn.Parameter(torch.empty(1, 1, out_channels))
        self.att_dst = nn.Parameter(torch.empty(1, 1, out_channels))
        self.bias = nn.Parameter(torch.zeros(out_channels))
        _glorot(self.lin_src.weight)
        _glorot(self.att_src)
        _glorot(self.att_dst)

    def forward(self, x, edge_index):
        n = x.size(0)
        h = self.lin_src(x)
        a_s = (h * self.att_src.view(1, -1)).sum(-1)
        a_d = (h * self.att_dst.view(1, -1)).sum(-1)
        s, t = edge_index[0], edge_index[1]
        z = torch.nn.functional.leaky_relu(a_s[s] + a_d[t], self.negative_slope)
        zmax = torch.full((n,), float("-inf"), dtype=z.dtype)
        zmax = zmax.scatter_reduce(0, t, z, reduce="amax", include_self=True)
        ez = torch.exp(z - zmax[t])
        den = torch.zeros(n, dtype=z.dtype).index_add_(0, t, ez)
        alpha = ez / (den[t] + 1e-16)
        out = torch.zeros_like(h).index_add_(0, t, alpha.unsqueeze(1) * h[s])
        return out + self.bias


# --------------------------------------------------------------------------- MessagePassing
class _Inspector:
    def __init__(self, owner):
        self.owner = owner

    def params(self, name, pop_first=False):
        p = list(inspect.signature(getattr(self.owner, name)).parameters)
        return p[1:] if pop_first else p

    def distribute(self, name, d):
        return {k: d[k] for k in self.params(name, pop_first=(name != "message")) if k in d}


class MessagePassing(nn.Module):
    """Only what the reference's overridden `propagate` uses (A.3)."""
    special_args = {"edge_index", "adj_t", "edge_index_i", "edge_index_j", "size", "size_i",
                    "size_j", "ptr", "index", "dim_size"}

    def __init__(self, aggr="add", flow="source_to_target", node_dim=-2):
        super().__init__()
        assert flow == "source_to_target"
        self.aggr, self.flow, self.node_dim = aggr, flow, node_dim
        self.inspector = _Inspector(self)
        self.fuse = False
        self.__explain__ = False
        user = set(self.inspector.params("message")) | set(self.inspector.params("aggregate", True)) \
            | set(self.inspector.params("update", True))
        self.__user_args__ = user - self.special_args
        self.__fused_user_args__ = set()

    def __check_input__(self, edge_index, size):
        assert isinstance(edge_index, torch.Tensor) and edge_index.dtype == torch.long
        assert edge_index.dim() == 2 and edge_index.size(0) == 2
        the_size = [None, None]
        if size is not None:
            the_size[0], the_size[1] = size[0], size[1]
        return the_size

    def __collect__(self, args, edge_index, size, kwargs):
        i, j = 1, 0  # source_to_target
        out = {}
        for arg in args:
            if arg[-2:] not in ("_i", "_j"):
                out[arg] = kwargs.get(arg, inspect.Parameter.empty)
            else:
                dim = j if arg[-2:] == "_j" else i
                data = kwargs.get(arg[:-2], inspect.Parameter.empty)
                if isinstance(data, torch.Tensor):
                    data = data.index_select(self.node_dim, edge_index[dim])
                out[arg] = data
        out["edge_index"] = edge_index
        out["index"] = edge_index[i]
        out["size"] = size
        out["dim_size"] = size[1]
        return out


# --------------------------------------------------------------------------- install
def _mod(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def install():
    """Pre-populate sys.modules with the stand-ins (idempotent)."""
    if "torch_geometric" in sys.modules and getattr(sys.modules["torch_geometric"], "_b3d_shim", False):
        return
    from typing import Optional, Tuple
    typing_mod = _mod("torch_geometric.typing", Adj=torch.Tensor, Size=Optional[Tuple[int, int]])
    nn_mod = _mod("torch_geometric.nn", MessagePassing=MessagePassing, GATConv=GATConv,
                  knn_graph=knn_graph, Sequential=nn.Sequential)
    tg = _mod("torch_geometric", nn=nn_mod, typing=typing_mod, _b3d_shim=True)
    tg.__path__ = []
    _mod("torch_scatter", scatter=scatter, gather_csr=None, segment_csr=None)
    _mod("torch_sparse", SparseTensor=type("SparseTensor", (), {}))
    mpl = _mod("matplotlib", pyplot=None)
    mpl.__path__ = []
    mpl.pyplot = _mod("matplotlib.pyplot")


def load_reference(root="/root/reference"):
    """Import the unmodified reference model modules. Build-container only."""
    import importlib
    import os
    if not os.path.isdir(root):
        raise RuntimeError(f"reference tree {root} not present (it never is on the GPU box)")
    install()
    # A product-side `batch_3dmot` re-export package may already be imported; drop it.
    for k in [k for k in sys.modules if k == "batch_3dmot" or k.startswith("batch_3dmot.")]:
        del sys.modules[k]
    # The reference's batch_3dmot/ has no __init__.py (namespace package), so a regular package of
    # the same name elsewhere on sys.path would shadow it: pin the package path explicitly.
    top = types.ModuleType("batch_3dmot")
    top.__path__ = [os.path.join(root, "batch_3dmot")]
    sys.modules["batch_3dmot"] = top
    pkg = importlib.import_module("batch_3dmot.models")
    for missing in ("heterolinear", "message_passing", "attention_message_passing"):
        m = _mod(f"batch_3dmot.models.{missing}", HeteroLinear=None, Linear=None)
        setattr(pkg, missing, m)
    pose = importlib.import_module("batch_3dmot.models.pose_gnn")
    clr = importlib.import_module("batch_3dmot.models.clr_att_gnn")
    # leave no trace of the reference package for later product imports
    for k in [k for k in sys.modules if k == "batch_3dmot" or k.startswith("batch_3dmot.")]:
        del sys.modules[k]
    return pose, clr


# --------------------------------------------------------------------------- graph-construction utilities
class _Quaternion:
    """Stand-in for pyquaternion.Quaternion restricted to what geo_utils.quaternion_yaw touches: a rotation
    about z by `yaw` and its `rotation_matrix`."""

    def __init__(self, yaw):
        self.yaw = float(yaw)

    @property
    def rotation_matrix(self):
        import numpy as np
        c, s_ = np.cos(self.yaw), np.sin(self.yaw)
        return np.array([[c, -s_, 0.0], [s_, c, 0.0], [0.0, 0.0, 1.0]])


class _Box:
    """Stand-in for nuscenes.utils.data_classes.Box: the attributes geo_utils / graph_utils read."""

    def __init__(self, center, wlh, yaw, velocity, name="car", token=None, score=1.0):
        import numpy as np
        self.center, self.wlh, self.velocity = np.asarray(center, float), np.asarray(wlh, float), np.asarray(velocity, float)
        self.orientation, self.name, self.token, self.score = _Quaternion(yaw), name, token, score


def load_reference_graph_utils(root="/root/reference"):
    """Import the unmodified batch_3dmot/utils/{geo_utils,graph_utils}.py (build container only) with stand-ins
    for pyquaternion / nuscenes / shapely. Returns (geo_utils, graph_utils, Box)."""
    import importlib
    import os
    if not os.path.isdir(root):
        raise RuntimeError(f"reference tree {root} not present (it never is on the GPU box)")
    _mod("pyquaternion", Quaternion=_Quaternion)
    ns = _mod("nuscenes"); ns.__path__ = []
    nu = _mod("nuscenes.utils"); nu.__path__ = []
    _mod("nuscenes.utils.data_classes", Box=_Box)
    sh = _mod("shapely"); sh.__path__ = []
    _mod("shapely.geometry", Polygon=type("Polygon", (), {}))
    for k in [k for k in sys.modules if k == "batch_3dmot" or k.startswith("batch_3dmot.")]:
        del sys.modules[k]
    top = types.ModuleType("batch_3dmot")
    top.__path__ = [os.path.join(root, "batch_3dmot")]
    sys.modules["batch_3dmot"] = top
    utils = types.ModuleType("batch_3dmot.utils")           # skip the package's other modules (nuScenes devkit imports)
    utils.__path__ = [os.path.join(root, "batch_3dmot", "utils")]
    sys.modules["batch_3dmot.utils"] = utils
    geo = importlib.import_module("batch_3dmot.utils.geo_utils")
    gu = importlib.import_module("batch_3dmot.utils.graph_utils")
    for k in [k for k in sys.modules if k == "batch_3dmot" or k.startswith("batch_3dmot.")]:
        del sys.modules[k]
    return geo, gu, _Box


def load_reference_graph_dataset(root="/root/reference"):
    """Import the unmodified batch_3dmot/utils/graph_data.py (build container only) and return its GraphDataset
    class. Stand-ins: torch_geometric.data.{Dataset, Data} (attribute bags), torch_geometric.{utils, transforms},
    batch_3dmot.predict_contrastive (an unused import), and batch_3dmot.utils.dataset reduced to get_class_config
    (dataset.py:33-51; the real module needs PIL and the nuScenes devkit)."""
    import importlib
    import os
    from types import SimpleNamespace
    if not os.path.isdir(root):
        raise RuntimeError(f"reference tree {root} not present (it never is on the GPU box)")
    install()

    class Dataset:
        def __init__(self, *a, **k):
            pass

    class Data(SimpleNamespace):
        pass

    tg = sys.modules["torch_geometric"]
    tg.data = _mod("torch_geometric.data", Dataset=Dataset, Data=Data)
    tg.utils = _mod("torch_geometric.utils")
    tg.transforms = _mod("torch_geometric.transforms")
    for k in [k for k in sys.modules if k == "batch_3dmot" or k.startswith("batch_3dmot.")]:
        del sys.modules[k]
    top = types.ModuleType("batch_3dmot")
    top.__path__ = [os.path.join(root, "batch_3dmot")]
    sys.modules["batch_3dmot"] = top
    utils = types.ModuleType("batch_3dmot.utils")
    utils.__path__ = [os.path.join(root, "batch_3dmot", "utils")]
    sys.modules["batch_3dmot.utils"] = utils
    top.utils = utils
    _mod("batch_3dmot.predict_contrastive", inference=None)

    def get_class_config(params, class_dict_name="nuscenes_tracking_eval"):        # dataset.py:33-51
        assert isinstance(class_dict_name, str)
        return vars(params.classes)[class_dict_name]

    utils.dataset = _mod("batch_3dmot.utils.dataset", get_class_config=get_class_config)
    gd = importlib.import_module("batch_3dmot.utils.graph_data")
    for k in [k for k in sys.modules if k == "batch_3dmot" or k.startswith("batch_3dmot.")]:
        del sys.modules[k]
    return gd.GraphDataset


def load_reference_predict_functions(root="/root/reference"):
    """The UNMODIFIED source text of predict.py's track-assembly functions (greedy_filter_node_flux,
    aggregate_node_flux, get_instance_metadata, combine_batches_to_scene, create_trajectories), cut out of
    /root/reference/batch_3dmot/predict.py by AST line ranges and exec'd in a namespace of stand-ins (the module
    itself cannot be imported: it parses argv, opens nuScenes and starts ray at import time). Returns the
    namespace dict; the caller sets ns['nusc'], ns['load_batch_detections'] and passes a stub gnn. Build
    container only."""
    import ast
    import json
    import os
    from collections import defaultdict
    from types import SimpleNamespace
    import numpy as np
    path = os.path.join(root, "batch_3dmot", "predict.py")
    if not os.path.exists(path):
        raise RuntimeError(f"reference tree {root} not present (it never is on the GPU box)")
    src = open(path).read()
    wanted = ("greedy_filter_node_flux", "aggregate_node_flux", "get_instance_metadata", "combine_batches_to_scene",
              "create_trajectories")

    def tqdm(it=None, *a, **k):
        return it
    tqdm.write = lambda *a, **k: None

    class Data(SimpleNamespace):
        def to(self, device):
            return self

    ns = {"np": np, "torch": torch, "os": os, "json": json, "defaultdict": defaultdict, "tqdm": tqdm, "ParamLib": object,
          "torch_geometric": SimpleNamespace(data=SimpleNamespace(Data=Data)),
          "batch_3dmot": SimpleNamespace(
              models=SimpleNamespace(cl_att_gnn=SimpleNamespace(GNN=object)),
              utils=SimpleNamespace(dataset=SimpleNamespace(
                  get_class_config=lambda params, class_dict_name: vars(params.classes)[class_dict_name])))}
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name in wanted:
            exec(compile(ast.get_source_segment(src, node), path, "exec"), ns)
    assert all(w in ns for w in wanted)
    return ns
