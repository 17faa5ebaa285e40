"""TMA row-gathered operands (b3d_linear_tma with `idx` segments, tile::gather4) and the edge block built on them:
kernel-level against a float64 reference on bf16-rounded operands, model-level against the per-node pre-projected
path (same function, different summation order) and the fp32 oracle."""
from types import SimpleNamespace

import pytest
import torch

from oracle import ref_restated as R
from batch3dmot_b200 import _lib as L, ops, synth
from batch3dmot_b200.clr_att_gnn import GNN
from batch3dmot_b200.pose_gnn import PoseGNN

pytestmark = pytest.mark.gpu
DEV = "cuda"


def bf(t):
    return t.to(torch.bfloat16).to(torch.float64)


def rel(a, b):
    b = b.double().cpu()
    return float((a.detach().double().cpu() - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.fixture(autouse=True)
def gather_mode():
    old = dict(ops.FEATURES)
    ops.FEATURES["gather_tma"] = True
    ops.set_precision("bf16")
    ops.invalidate_weight_cache()
    yield
    ops.FEATURES.update(old)
    ops.set_precision("fp32")
    ops.invalidate_weight_cache()


@pytest.mark.parametrize("M,layout,n_out", [(1000, ((96, 0), (96, 1), (64, None), (64, None)), 256),
                                            (70000, ((96, 0), (64, None), (96, 0)), 192),
                                            (5000, ((48, 1), (48, 0), (32, None)), 96),
                                            (300, ((64, None), (40, 1)), 64),
                                            (128 * 148 * 2 + 5, ((96, 1), (64, None), (96, 1)), 192)])
def test_linear_tma_row_gathered_segments(M, layout, n_out):
    """layout: (width, which index array gathers it | None = dense edge rows)."""
    torch.manual_seed(M + n_out)
    Nn = 777
    idx = [torch.randint(0, Nn, (M,)), torch.randint(0, Nn, (M,))]
    xs = [torch.randn(Nn if sel is not None else M, w).to(torch.bfloat16) for w, sel in layout]
    cat = torch.cat([x.double()[idx[sel]] if sel is not None else x.double() for x, (w, sel) in zip(xs, layout)], 1)
    idx_d = [i.int().to(DEV) for i in idx]
    items = [(x.to(DEV), idx_d[sel] if sel is not None else None, None, 0) for x, (w, sel) in zip(xs, layout)]
    K = sum(w for w, _ in layout)
    W, b = torch.randn(n_out, K) * 0.1, torch.randn(n_out)
    bits = ops.new_relu_bits(M, n_out, DEV) if n_out % 32 == 0 else None
    n0 = L.launch_count()
    y = ops.linear_raw(items, W.to(DEV), b.to(DEV), M, L.ACT_RELU, tc=True, out_dtype=torch.bfloat16, bits_out=bits)
    assert L.launch_count() - n0 == 2                     # segment-padded pack + the TMA kernel
    ref = torch.relu(cat @ bf(W).t() + b.double())
    assert rel(y, ref) < 1e-2                             # bf16 output rounding
    y32 = ops.linear_raw(items, W.to(DEV), b.to(DEV), M, L.ACT_RELU, tc=True, out_dtype=torch.float32)
    assert rel(y32, ref) < 1e-5                           # fp32 output: only the accumulation order differs


def _mm_kw(d):
    return dict(x_img=d.x_img, pointnet_out=d.pointnet_out, radarnet_out=d.radarnet_out, lidar_mask=d.m_lidar,
                radar_mask=d.m_radar)


def to_dev(ns):
    return SimpleNamespace(**{k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in vars(ns).items()})


@pytest.mark.parametrize("multimodal", [True, False])
def test_gathered_edge_block_matches_preprojected_path_and_oracle(multimodal):
    data = synth.scene_graph(seed=77, T=12, nodes_per_frame=40)
    if multimodal:
        data = synth.add_modalities(data, 77, raw=False)
    data = synth.add_labels(data, 77)
    d = to_dev(data)
    res = {}
    for on in (True, False):
        ops.FEATURES["gather_tma"] = on
        ops.FEATURES["edge_block"] = on
        ops.invalidate_weight_cache()
        torch.manual_seed(5621)
        m = (GNN(None, None, None) if multimodal else PoseGNN()).to(DEV)
        n0 = L.launch_count()
        out, _ = m(d, **_mm_kw(d)) if multimodal else m(d)
        loss = ops.bce_loss(out, d.y, d.edge_weights, batch_size=2, from_logits=not multimodal)
        loss.backward()
        res[on] = (out.detach().float().cpu(), {k: p.grad.detach().float().cpu() for k, p in m.named_parameters()
                                                if p.grad is not None}, L.launch_count() - n0)
        sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    with torch.no_grad():
        out_ref = (R.mm_gnn_forward(sd, data) if multimodal else R.pose_gnn_forward(sd, data))[0]
    assert rel(res[True][0], out_ref) < 2e-2               # north_star tolerance for bf16 tiles
    assert rel(res[True][0], res[False][0]) < 1e-2
    fro = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))
    for k, gv in res[False][1].items():
        assert fro(res[True][1][k], gv) < 5e-2, k
