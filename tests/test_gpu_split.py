"""The tensor-core path of the 1e-4 ("fp32") parity mode: split-bf16 tiles (hi + lo bf16 parts of every fp32 operand,
three tcgen05.mma per K step, fp32 accumulation in TMEM) against a float64 reference, and the drop-in models in that
mode against the golden vectors of the unmodified reference at the tolerance north_star states for fp32 (1e-4)."""
from types import SimpleNamespace

import pytest
import torch

from batch3dmot_b200 import _lib as L, ops
from batch3dmot_b200.pose_gnn import PoseGNN
from batch3dmot_b200.clr_att_gnn import GNN
from .conftest import load_golden
from .test_gpu_models import check_grads

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ops.FEATURES["split_tc"], reason="split-bf16 tiles not switched on")]
DEV = "cuda"


def rel(a, b):
    b = b.double().cpu()
    return float((a.detach().double().cpu() - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.fixture(autouse=True)
def fp32_mode():
    ops.set_precision("fp32")
    ops.invalidate_weight_cache()
    yield
    ops.set_precision("fp32")


@pytest.mark.parametrize("M,widths,n_out", [(1024, (64,), 64), (3000, (48, 48, 32), 96), (5000, (96, 96, 64, 64), 256),
                                            (2000, (64, 128, 96, 64, 128, 96, 64), 512), (700, (128,), 48)])
def test_split_linear_and_wgrad_vs_float64(M, widths, n_out):
    torch.manual_seed(M + n_out)
    N = 777
    xs = [torch.randn(N if s % 2 == 0 else M, w) for s, w in enumerate(widths)]
    idx = torch.randint(0, N, (len(widths), M))
    cat = torch.cat([x[idx[s]] if x.size(0) == N else x for s, x in enumerate(xs)], 1).double()
    items = [(x.to(DEV), idx[s].int().to(DEV) if x.size(0) == N else None, None, 0) for s, x in enumerate(xs)]
    K = sum(widths)
    W, b = torch.randn(n_out, K) * 0.1, torch.randn(n_out)
    n0 = L.launch_count()
    y = ops.linear_raw(items, W.to(DEV), b.to(DEV), M, L.ACT_RELU)
    assert L.launch_count() - n0 == 2                       # split pack + tcgen05 tile kernel (not the FFMA kernel)
    assert y.dtype == torch.float32
    ref = torch.relu(cat @ W.double().t() + b.double())
    assert rel(y, ref) < 2e-5                                # ~2^-16 per product, against 4e-3 for plain bf16 tiles
    # input gradient with an fp32 ReLU mask + accumulate
    dy, hmask = torch.randn(M, n_out), torch.randn(M, K)
    out = torch.full((M, K), 0.25, device=DEV)
    ops.linear_raw([(dy.to(DEV), None, None, 0)], W.to(DEV), None, M, trans_w=True, out=out, accumulate=True,
                   out_mask=hmask.to(DEV))
    assert rel(out, 0.25 + (dy.double() @ W.double()) * (hmask > 0)) < 2e-5
    # weight gradient (deterministic row split) + bias gradient
    dW, db = ops.wgrad_raw((dy.to(DEV), None, None, 0), items, M, n_out, K)
    assert rel(dW, dy.double().t() @ cat) < 2e-5 and rel(db, dy.double().sum(0)) < 2e-5
    dW2, _ = ops.wgrad_raw((dy.to(DEV), None, None, 0), items, M, n_out, K)
    assert torch.equal(dW, dW2)


def _run(model_cls, golden, loss_kw, multimodal):
    g = load_golden(golden)
    d = SimpleNamespace(**{k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in g["data"].items()})
    m = (GNN(None, None, None) if multimodal else PoseGNN()).to(DEV)
    m.load_state_dict(g["state_dict"])
    kw = dict(x_img=d.x_img, pointnet_out=d.pointnet_out, radarnet_out=d.radarnet_out, lidar_mask=d.m_lidar,
              radar_mask=d.m_radar) if multimodal else {}
    out, _ = m(d, **kw)
    loss = ops.bce_loss(out, d.y, d.edge_weights, **loss_kw)
    loss.backward()
    assert rel(out, g["out"]) < 1e-4 and abs(loss.item() - g["loss"].item()) < 1e-5
    check_grads(m, g["grads"])


def test_models_in_split_mode_match_the_reference_golden_vectors():
    """Same assertions as tests/test_gpu_models.py, and a check that the tensor-core kernels really ran."""
    n0 = L.launch_count()
    _run(PoseGNN, "pose_small.pt", dict(from_logits=True), False)
    _run(GNN, "mm_small.pt", dict(batch_size=2), True)
    assert L.launch_count() > n0


def test_split_mode_matches_exact_mode_on_a_scene():
    from batch3dmot_b200 import synth
    data = synth.add_labels(synth.add_modalities(synth.scene_graph(seed=5621), 5621, raw=False), 5621)
    d = SimpleNamespace(**{k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in vars(data).items()})
    kw = dict(x_img=d.x_img, pointnet_out=d.pointnet_out, radarnet_out=d.radarnet_out, lidar_mask=d.m_lidar,
              radar_mask=d.m_radar)
    res = {}
    for mode in ("exact", "fp32"):
        ops.set_precision(mode)
        torch.manual_seed(5621)
        m = GNN(None, None, None).to(DEV)
        out, _ = m(d, **kw)
        ops.bce_loss(out, d.y, d.edge_weights, batch_size=2).backward()
        res[mode] = (out.detach(), {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None})
    assert rel(res["fp32"][0], res["exact"][0].cpu()) < 1e-4
    for k, gv in res["exact"][1].items():
        assert rel(res["fp32"][1][k], gv.cpu()) < 1e-3, k
