"""CPU tests: the oracle restatement against the committed golden vectors (produced by the
unmodified reference under oracle/pyg_shim.py, see oracle/gen_golden.py) and, when the
reference tree is present (build container), against the live reference."""
import os
from types import SimpleNamespace

import pytest
import torch

from oracle import ref_restated as R
from batch3dmot_b200 import synth
from .conftest import load_golden


def _data(g):
    return SimpleNamespace(**g["data"])


def _check_grads(params, packed, rtol=2e-6, skip_qk=False):
    for k, ref in packed.items():
        got = params[k].grad
        assert got is not None, k
        if isinstance(ref, dict):
            samp = got.reshape(-1)[::ref["stride"]]
            scale = max(1.0, float(ref["sample"].abs().max()))
            if skip_qk and "in_proj" in k:
                n = got.size(0) // 3 * 2 * (got.numel() // got.size(0))
                mask = (torch.arange(got.numel())[::ref["stride"]] >= n)
                assert (samp - ref["sample"])[mask].abs().max() <= rtol * scale, k
                continue
            assert (samp - ref["sample"]).abs().max() <= rtol * scale, k
            assert abs(float(got.double().sum()) - float(ref["sum"])) <= 1e-4 * max(1.0, float(ref["abs_sum"])), k
        else:
            if skip_qk and "in_proj" in k:
                D = ref.size(0) // 3
                got, ref = got[2 * D:], ref[2 * D:]
            assert (got - ref).abs().max() <= rtol * max(1.0, float(ref.abs().max())), k


def test_pose_restatement_matches_golden():
    g = load_golden("pose_small.pt")
    data = _data(g)
    params = {k: v.clone().requires_grad_(True) for k, v in g["state_dict"].items()}
    out, x_enc = R.pose_gnn_forward(params, data)
    assert torch.equal(out, g["out"]) and torch.equal(x_enc, g["x_enc"])      # bit-exact, CPU fp32
    loss = R.bce_logits_loss(out, data.y, data.edge_weights)
    assert abs(loss.item() - g["loss"].item()) <= 1e-7
    loss.backward()
    _check_grads(params, g["grads"])
    assert all(params[k].grad is None for k in params if k.startswith("knn_conv"))  # C1: dead k-NN conv
    # one MP iteration in isolation
    with torch.no_grad():
        e = R.mlp(params, "edge_encoder", data.edge_attr.float(), (0, 2, 4))
        x0 = R.mlp(params, "node_encoder", data.pose_feats, (0, 2, 4))
        x1, e1 = R.causal_mp(params, x0, data.edge_index, e, x0)
    assert torch.equal(x1, g["mp1_x"]) and torch.equal(e1, g["mp1_e"])


def test_mm_restatement_matches_golden():
    g = load_golden("mm_small.pt")
    data = _data(g)
    params = {k: v.clone().requires_grad_(True) for k, v in g["state_dict"].items()}
    out, x_sens = R.mm_gnn_forward(params, data)
    assert torch.equal(out, g["out"]) and torch.equal(x_sens, g["x_sens"])
    assert float(out.min()) > 0 and float(out.max()) < 1
    loss = R.bce_loss(out, data.y, data.edge_weights, batch_size=2)
    assert abs(loss.item() - g["loss"].item()) <= 1e-7
    loss.backward()
    _check_grads(params, g["grads"], skip_qk=True)


def test_mha_len1_closed_form_is_exact():
    """C2: nn.MultiheadAttention with one key == out_proj(W_v v + b_v), train and eval."""
    torch.manual_seed(0)
    for D in (96, 128, 64):
        att = torch.nn.MultiheadAttention(D, 2, kdim=D, vdim=D, batch_first=True)
        p = {f"a.{k}": v for k, v in att.state_dict().items()}
        v, q = torch.randn(37, D), torch.randn(37, D)
        v[5] = 0
        ref, _ = att(query=q.unsqueeze(1), key=v.unsqueeze(1), value=v.unsqueeze(1), need_weights=False)
        assert (ref.squeeze(1) - R.mha_len1(p, "a", v)).abs().max().item() == 0.0


def test_knn_gat_golden():
    g = load_golden("knn_gat_small.pt")
    idx = R.knn_frames(g["x"], g["frame_ptr"], g["k"])
    assert torch.equal(idx, g["idx"])
    y = R.gat_conv(g["gat_state_dict"], g["x"], R.knn_to_edge_index(idx))
    assert torch.equal(y, g["gat_out"])


def test_knn_edge_cases():
    x = torch.tensor([[0.0], [1.0], [1.0], [3.0], [9.0]])
    ptr = torch.tensor([0, 4, 5])                     # frame of 4, frame of 1
    idx = R.knn_frames(x, ptr, 20)
    assert idx[4].tolist() == [-1] * 20               # n_t = 1 -> no neighbours
    assert idx[0, :3].tolist() == [1, 2, 3] and idx[0, 3] == -1   # k_eff = n_t - 1, tie by index
    assert idx[1, :3].tolist() == [2, 0, 3]
    ei = R.knn_to_edge_index(idx)
    assert ei.size(1) == 12 and bool((ei[0] != ei[1]).all())


def test_csr_spec():
    torch.manual_seed(1)
    idx = torch.randint(0, 50, (1000,))
    rowptr, perm = R.csr_build(idx, 64)
    assert rowptr[-1] == 1000 and torch.equal(idx[perm], idx.sort(stable=True).values)
    for n in (0, 7, 49, 63):
        seg = perm[rowptr[n]:rowptr[n + 1]]
        assert bool((idx[seg] == n).all()) and bool((seg[1:] > seg[:-1]).all())   # stable


def test_faithful_mode_is_identical():
    """The op-faithful variant (dead k-NN + node encoder twice) used as the CPU baseline gives the
    same outputs as the lean restatement."""
    g = load_golden("pose_small.pt")
    with torch.no_grad():
        a, _ = R.pose_gnn_forward(g["state_dict"], _data(g), faithful=True)
    assert torch.equal(a, g["out"])


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree only exists in the build container")
def test_live_reference_matches_restatement():
    from oracle import pyg_shim
    pose_mod, clr_mod = pyg_shim.load_reference()
    torch.manual_seed(11)
    ref = pose_mod.PoseGNN()
    g = synth.scene_graph(seed=3, T=6, nodes_per_frame=9, k=6)
    with torch.no_grad():
        out, x_enc = ref(g)
        o2, x2 = R.pose_gnn_forward(dict(ref.state_dict()), g)
    assert torch.equal(out, o2) and torch.equal(x_enc, x2)
    synth.add_modalities(g, seed=3)
    enc = [synth.EmbeddingEncoder(t) for t in (g.x_img, g.pointnet_out, g.radarnet_out)]
    ref = clr_mod.GNN(*enc)
    with torch.no_grad():
        out, xs = ref(g)
        o2, xs2 = R.mm_gnn_forward(dict(ref.state_dict()), g)
    assert torch.equal(out, o2) and torch.equal(xs, xs2)
