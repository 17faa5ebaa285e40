"""CPU tests of the host logic: C-ABI exports, drop-in module surface (constructor args,
state_dict keys / shapes / default-init parity with the reference), generators."""
import ctypes
import os
import re

import pytest
import torch

from batch3dmot_b200 import _lib, synth
from .conftest import ROOT, load_golden


def test_library_exports_every_declared_symbol(lib_built):
    hdr = open(os.path.join(ROOT, "include", "b3d.h")).read()
    declared = sorted(set(re.findall(r"\b(b3d_[a-z0-9_]+)\s*\(", hdr)))
    assert declared == _lib.exported_symbols(), "ctypes table and include/b3d.h disagree"
    so = ctypes.CDLL(lib_built)
    for name in declared:
        assert hasattr(so, name), f"{name} not exported by libb3d.so"
    # host-only entry points are callable without a GPU
    so.b3d_csr_workspace_bytes.restype = ctypes.c_size_t
    so.b3d_csr_workspace_bytes.argtypes = [ctypes.c_int64, ctypes.c_int64]
    assert so.b3d_csr_workspace_bytes(1000, 100) > 0
    so.b3d_last_error.restype = ctypes.c_char_p
    assert isinstance(so.b3d_last_error(), bytes)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.lib()


def test_pose_gnn_dropin_surface():
    from batch_3dmot.models.pose_gnn import PoseGNN, CausalMessagePassing   # reference import path
    g = load_golden("pose_small.pt")
    torch.manual_seed(5621)
    m = PoseGNN(gnn_depth=6, edge_dim=16, node_dim=19, mp_type="attention")
    sd = m.state_dict()
    ref = g["state_dict"]
    assert set(sd) == set(ref)                # incl. PyG 2.0.x's duplicate knn_conv.lin_dst.weight
    for k in sd:
        assert sd[k].shape == ref[k].shape, k
        assert torch.equal(sd[k], ref[k]), f"default init differs from the reference at {k}"
    m.load_state_dict(ref, strict=True)       # PyG 2.0.x checkpoints carry lin_dst.weight too
    assert isinstance(m.message_passing, CausalMessagePassing)
    assert sum(p.numel() for p in m.parameters()) == 86605


def test_mm_gnn_dropin_surface():
    from batch_3dmot.models.clr_att_gnn import GNN
    g = load_golden("mm_small.pt")
    d = g["data"]
    enc = [synth.EmbeddingEncoder(d[k]) for k in ("x_img", "pointnet_out", "radarnet_out")]
    torch.manual_seed(5621)
    m = GNN(*enc, use_attention=True, gnn_depth=6, edge_dim=64, node_dim=179)
    sd, ref = m.state_dict(), g["state_dict"]
    assert set(sd) == set(ref)                # incl. PyG 2.0.x's duplicate knn_conv.lin_dst.weight
    for k in sd:
        assert torch.equal(sd[k], ref[k]), f"default init differs from the reference at {k}"
    m.load_state_dict(ref, strict=True)
    trainable = sum(p.numel() for n, p in m.named_parameters() if p.requires_grad)
    assert trainable == 1319697
    with pytest.raises(NotImplementedError):
        GNN(*enc, use_attention=False)


def test_newer_pyg_gat_key_spelling_loads():
    from batch3dmot_b200.gat import GATConv
    c = GATConv(48, 48)
    sd = c.state_dict()
    sd.pop("lin_dst.weight")                  # newer PyG: one `lin.weight` instead of lin_src / lin_dst
    sd["lin.weight"] = sd.pop("lin_src.weight") + 1
    c.load_state_dict(sd, strict=True)
    assert torch.equal(c.lin_src.weight, sd["lin.weight"])


def test_scene_graph_properties():
    g = synth.scene_graph(seed=5621, T=10, nodes_per_frame=20, k=40)
    src, dst = g.edge_index
    assert g.pose_feats.shape == (200, 19) and g.edge_attr.dtype == torch.float64
    assert bool((src < dst).all()) and bool((dst[1:] >= dst[:-1]).all())
    assert bool((g.node_classes[src] == g.node_classes[dst]).all())          # category-disjoint
    dt = g.node_timestamps[dst] - g.node_timestamps[src]
    assert int(dt.min()) >= 1 and int(dt.max()) <= 4
    assert torch.equal(g.edge_attr[:, 3], dt.double())
    assert int(torch.bincount(dst).max()) <= 40
    g2 = synth.scene_graph(seed=5621, T=10, nodes_per_frame=20, k=40)
    assert torch.equal(g.edge_index, g2.edge_index) and torch.equal(g.pose_feats, g2.pose_feats)


def test_config1_shape():
    g = synth.scene_graph(seed=5621)          # T=40, 50 nodes/frame
    assert g.num_nodes == 2000 and 50000 < g.edge_index.size(1) < 70000


def test_collate_and_windows():
    a = synth.add_labels(synth.add_modalities(synth.scene_graph(seed=1, T=7, nodes_per_frame=6, k=5), 1), 1)
    b = synth.add_labels(synth.add_modalities(synth.scene_graph(seed=2, T=6, nodes_per_frame=4, k=5), 2), 2)
    c = synth.collate([a, b])
    assert c.num_nodes == a.num_nodes + b.num_nodes
    assert torch.equal(c.edge_index[:, a.edge_index.size(1):], b.edge_index + a.num_nodes)
    assert torch.equal(c.batch, torch.cat([torch.zeros(a.num_nodes), torch.ones(b.num_nodes)]).long())
    assert c.lidar_feats.shape == (c.num_nodes, 128, 3)
    ws = synth.windows(a, 5)
    assert len(ws) == 3
    for w in ws:
        assert int(w.node_timestamps.max() - w.node_timestamps.min()) <= 4
        s, d = w.edge_index
        assert bool((a.edge_index[:, w.global_edge_id] == torch.stack([w.global_node_id[s], w.global_node_id[d]])).all())


def test_cb_weights_match_reference_formula():
    # graph_data.py:126-138 evaluated by hand for 'car' (class id 3)
    beta, n = 0.8, 5 * synth.REL_FREQ_TRAIN["car"]
    w = synth.cb_weights(torch.tensor([3.0]))
    assert abs(float(w) - (1 - beta) / (1 - beta ** n)) < 1e-7


def test_split_cols_and_fanout_host_logic():
    """ops.split_cols: column slices whose backward is one concatenation (zero blocks for unused slices);
    ops.fanout degenerates to plain aliases without autograd. Pure host logic, no kernel involved."""
    from batch3dmot_b200 import ops
    torch.manual_seed(0)
    x = torch.randn(7, 12, requires_grad=True)
    a, b, c = ops.split_cols(x, (3, 4, 5))
    assert a.shape == (7, 3) and b.shape == (7, 4) and c.shape == (7, 5)
    assert torch.equal(torch.cat([a, b, c], 1), x)
    ga, gc = torch.randn(7, 3), torch.randn(7, 5)
    (a * ga).sum().add((c * gc).sum()).backward()          # slice b unused: its gradient block is zero
    assert torch.equal(x.grad, torch.cat([ga, torch.zeros(7, 4), gc], 1))
    y = torch.randn(4, 8)
    u, v = ops.fanout(y, 2)
    assert u is y and v is y
    with torch.no_grad():
        p, q = ops.split_cols(x, (6, 6))
        assert p.shape == (7, 6) and q.shape == (7, 6)


def test_focal_loss_oracle_reduces_to_bce():
    """oracle.ref_restated.focal_loss (published definition; parity unpinned, not in the reference) with gamma = 0,
    alpha = 0.5 is half the reference's BCELoss; gamma > 0 down-weights easy examples."""
    from oracle import ref_restated as R
    torch.manual_seed(1)
    p = torch.rand(500) * 0.98 + 0.01
    y = (torch.rand(500) < 0.2).long()
    w = torch.rand(500) + 0.5
    assert abs(float(R.focal_loss(p, y, w, gamma=0.0, alpha=0.5)) - 0.5 * float(R.bce_loss(p, y, w))) < 1e-6
    assert float(R.focal_loss(p, y, w, gamma=2.0, alpha=0.5)) < float(R.focal_loss(p, y, w, gamma=0.0, alpha=0.5))


def test_grad_target_windows_alias_the_parameter_gradient():
    """ops._grad_target (CPU: pure tensor bookkeeping): inside ops.grads_into_params() a leaf parameter maps to its .grad,
    a row / column slice view of it to the same window of .grad (also when the parameter and its gradient are views of
    flat buffers, as parallel.FlatParams lays them out); everything else maps to None."""
    import torch
    from batch3dmot_b200 import ops
    flat, gflat = torch.zeros(100 + 6 * 10), torch.zeros(100 + 6 * 10)
    p = torch.nn.Parameter(torch.randn(6, 10))
    p.data = flat[100:160].view(6, 10)
    assert ops._grad_target(p) is None                      # switch off
    with ops.grads_into_params():
        assert ops._grad_target(p) is None                  # no .grad yet
        p.grad = gflat[100:160].view(6, 10)
        assert ops._grad_target(p) is p.grad
        for view in (p[:, 3:7], p[2:5], p[1:4, 2:9], p[:, 4:]):
            t = ops._grad_target(view)
            assert t.shape == view.shape and t.stride() == view.stride()
            t.fill_(1.0)
            mark = torch.zeros_like(p.data)
            torch.as_strided(mark, view.size(), view.stride(), view.storage_offset() - p.storage_offset()).fill_(1.0)
            assert torch.equal(p.grad, mark)
            p.grad.zero_()
        assert ops._grad_target(p.detach().clone()) is None                    # not a leaf that requires grad
        assert ops._grad_target(torch.cat([p[:, :2], p[:, 5:]], 1)) is None      # derived (non-leaf, non-view)
        assert ops._grad_target(p.double()) is None
        q = torch.nn.Parameter(torch.randn(4, 4))
        q.grad = torch.zeros(4, 4).t()                                          # gradient laid out differently
        assert ops._grad_target(q) is None
    assert not ops._GRAD_SINK


def test_bucketed_trainer_size_ladder_and_padding_layout():
    """parallel._bucket / BucketedTrainer._static / _fill on CPU tensors: sizes land on a ladder with at most 12.5 %
    padding (plus the floor), padded rows are zero, dummy edges sit on the last dummy node with weight 0, dummy nodes get
    a frame of their own, and a smaller batch re-using the buffers leaves nothing of the previous one behind."""
    import torch
    from types import SimpleNamespace
    from batch3dmot_b200.parallel import BucketedTrainer, _bucket
    for n in (1, 63, 64, 65, 500, 513, 9629, 61561, 146961, 3925343):
        b = _bucket(n, 512)
        assert b >= n and b % 512 == 0 and (b - n) <= max(512, n // 8) and _bucket(b, 512) == b
    assert len({_bucket(n, 512) for n in range(8193, 16385)}) == 8          # 8 sizes per octave

    def batch(n, e, seed):
        g = torch.Generator().manual_seed(seed)
        return SimpleNamespace(pose_feats=torch.randn(n, 19, generator=g), node_timestamps=torch.randint(0, 5, (n,), generator=g),
                               edge_attr=torch.randn(e, 4, generator=g).double(), y=torch.randint(0, 2, (e,), generator=g),
                               edge_weights=torch.rand(e, generator=g), edge_index=torch.randint(0, n, (2, e), generator=g))
    d1, d2 = batch(100, 700, 1), batch(90, 650, 2)
    kw1 = {"x_img": torch.randn(100, 96), "lidar_mask": torch.ones(100, dtype=torch.bool)}
    kw2 = {"x_img": torch.randn(90, 96), "lidar_mask": torch.ones(90, dtype=torch.bool)}
    bt = BucketedTrainer.__new__(BucketedTrainer)
    n_pad, e_pad = _bucket(101, 64), _bucket(700, 512)
    assert (n_pad, e_pad) == (_bucket(91, 64), _bucket(650, 512)) == (128, 1024)
    s, skw = bt._static(d1, kw1, n_pad, e_pad)
    for d, kw, n, e in ((d1, kw1, 100, 700), (d2, kw2, 90, 650)):
        BucketedTrainer._fill(s, skw, d, kw, n, e)
        assert torch.equal(s.pose_feats[:n], d.pose_feats) and float(s.pose_feats[n:].abs().sum()) == 0
        assert torch.equal(s.edge_index[:, :e], d.edge_index) and bool((s.edge_index[:, e:] == n_pad - 1).all())
        assert torch.equal(s.edge_weights[:e], d.edge_weights) and float(s.edge_weights[e:].abs().sum()) == 0
        assert torch.equal(s.edge_attr[:e], d.edge_attr) and s.edge_attr.dtype == torch.float64
        assert torch.equal(s.y[:e], d.y) and int(s.y[e:].abs().sum()) == 0
        assert torch.equal(s.node_timestamps[:n], d.node_timestamps) and bool((s.node_timestamps[n:] > 10 ** 6).all())
        assert torch.equal(skw["x_img"][:n], kw["x_img"]) and float(skw["x_img"][n:].abs().sum()) == 0
        assert bool(skw["lidar_mask"][:n].all()) and not bool(skw["lidar_mask"][n:].any())
