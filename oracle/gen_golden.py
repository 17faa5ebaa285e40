"""Generate tests/golden/*.pt by running the UNMODIFIED reference model files
(/root/reference/batch_3dmot/models/{pose_gnn,clr_att_gnn}.py) on CPU under
oracle/pyg_shim.py. Build-container only (the reference tree does not travel).

    python -m oracle.gen_golden

Also asserts the pin: reference-under-shim == oracle/ref_restated.py with
max |diff| = 0.0 (forward) before writing anything.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
import os
import sys
from types import SimpleNamespace

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import pyg_shim, ref_restated as R  # noqa: E402
from batch3dmot_b200 import synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
SEED = 5621  # pose_config.yaml:96


def small_graph(seed):
    # 8 frames x 12 nodes, k=10: N=96, E ~ 1.5k -> fixtures stay small
    g = synth.scene_graph(seed=seed, T=8, nodes_per_frame=12, window=4, k=10)
    synth.add_modalities(g, seed=seed)
    synth.add_labels(g, seed=seed, p_pos=0.1)
    return g


def slim(g):
    """Fixture payload: drop the raw lidar/radar tensors (rebuilt from the masks by
    synth.add_raw_feats) to keep the files small."""
    return {k: v for k, v in vars(g).items() if k not in ("lidar_feats", "radar_feats", "img_feats")}


def pack_grads(gr, full_below=20000):
    """Full gradient for small parameters; float64 fingerprints (sum, |sum|, strided sample)
    for the big ones."""
    out = {}
    for k, v in gr.items():
        if v.numel() <= full_below:
            out[k] = v
        else:
            f = v.reshape(-1).double()
            out[k] = {"sum": f.sum(), "abs_sum": f.abs().sum(), "stride": 97, "sample": v.reshape(-1)[::97].clone()}
    return out


def grads_of(module, loss):
    module.zero_grad()
    loss.backward()
    return {k: p.grad.clone() for k, p in module.named_parameters() if p.grad is not None}


def main():
    os.makedirs(GOLD, exist_ok=True)
    pose_mod, clr_mod = pyg_shim.load_reference()
    torch.set_num_threads(1)

    # ---------------------------------------------------------------- PoseGNN
    torch.manual_seed(SEED)
    ref = pose_mod.PoseGNN()
    g = small_graph(SEED)
    out, x_enc = ref(g)
    sd = {k: v.detach().clone() for k, v in ref.state_dict().items()}
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    o2, x2 = R.pose_gnn_forward(params, g)
    assert (out - o2).abs().max().item() == 0.0 and torch.equal(x_enc, x2), "restatement != reference (pose)"
    o3, _ = R.pose_gnn_forward(params, g, faithful=True)
    assert torch.equal(out, o3)
    # gradient oracle: BCE-with-logits (C11), class-balanced weights
    loss = R.bce_logits_loss(out, g.y, g.edge_weights)
    gref = grads_of(ref, loss)
    loss2 = R.bce_logits_loss(o2, g.y, g.edge_weights)
    loss2.backward()
    for k, v in gref.items():
        d = (params[k].grad - v).abs().max().item()
        assert d <= 1e-6 * max(1.0, v.abs().max().item()), (k, d)
    assert all(k.startswith("knn_conv") for k in sd if k not in gref), "unexpected grad-less params"
    # one message-passing iteration in isolation (intermediates)
    with torch.no_grad():
        e_enc = ref.edge_encoder(g.edge_attr.float())
        x0 = ref.node_encoder(g.pose_feats)
        x1, e1 = ref.message_passing.forward(x0, g.edge_index, e_enc, x0)
    torch.save({
        "state_dict": sd, "data": slim(g), "out": out.detach(), "x_enc": x_enc.detach(),
        "loss": loss.detach(), "grads": pack_grads(gref), "mp1_x": x1, "mp1_e": e1,
    }, os.path.join(GOLD, "pose_small.pt"))
    print("pose_small: N=%d E=%d loss=%.6f" % (g.num_nodes, g.edge_index.size(1), loss.item()))

    # ---------------------------------------------------------------- multimodal GNN
    torch.manual_seed(SEED)
    g = small_graph(SEED + 7)
    enc = (synth.EmbeddingEncoder(g.x_img), synth.EmbeddingEncoder(g.pointnet_out),
           synth.EmbeddingEncoder(g.radarnet_out))
    ref = clr_mod.GNN(*enc, use_attention=True)
    out, x_sens = ref(g)
    sd = {k: v.detach().clone() for k, v in ref.state_dict().items()}
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    o2, xs2 = R.mm_gnn_forward(params, g)
    dmax = (out - o2).abs().max().item()
    assert dmax == 0.0 and torch.equal(x_sens, xs2), f"restatement != reference (multimodal) {dmax}"
    mha = {k: getattr(ref, k) for k in ("c2c_att", "l2l_att", "r2r_att")}
    o3, _ = R.mm_gnn_forward(params, g, faithful=True, mha_modules=mha)
    assert (out - o3).abs().max().item() == 0.0
    loss = R.bce_loss(out, g.y, g.edge_weights, batch_size=2)   # train.py:136-141, cl_config batch_size 2
    gref = grads_of(ref, loss)
    loss2 = R.bce_loss(o2, g.y, g.edge_weights, batch_size=2)
    loss2.backward()
    for k, v in gref.items():
        if "in_proj" in k:
            continue   # W_q/W_k rows get exactly-zero / rounding-noise grads in the reference (C2)
        d = (params[k].grad - v).abs().max().item()
        assert d <= 2e-6 * max(1.0, v.abs().max().item()), (k, d)
    for k in ("c2c_att", "l2l_att", "r2r_att"):
        D = sd[f"{k}.out_proj.weight"].size(0)
        gv = gref[f"{k}.in_proj_weight"]
        assert gv[:2 * D].abs().max().item() <= 1e-12, "q/k projection grads should vanish"
        d = (params[f"{k}.in_proj_weight"].grad[2 * D:] - gv[2 * D:]).abs().max().item()
        assert d <= 2e-6 * max(1.0, gv.abs().max().item()), (k, d)
    torch.save({
        "state_dict": sd, "data": slim(g), "out": out.detach(), "x_sens": x_sens.detach(),
        "loss": loss.detach(), "grads": pack_grads(gref),
    }, os.path.join(GOLD, "mm_small.pt"))
    print("mm_small: N=%d E=%d loss=%.6f lidar=%.2f radar=%.2f" % (
        g.num_nodes, g.edge_index.size(1), loss.item(), g.m_lidar.float().mean(), g.m_radar.float().mean()))

    # ---------------------------------------------------------------- k-NN + GAT (shim == restatement)
    x, ptr = synth.knn_stress(SEED, N=300, frame=60, D=48)
    idx = R.knn_frames(x, ptr, 20)
    for f in range(ptr.numel() - 1):
        a, b = int(ptr[f]), int(ptr[f + 1])
        ei = pyg_shim.knn_graph(x[a:b], k=20) + a
        mine = R.knn_to_edge_index(idx[a:b])
        mine[1] += a
        assert torch.equal(ei, mine)
    torch.manual_seed(SEED)
    gat = pyg_shim.GATConv(48, 48, add_self_loops=False)
    with torch.no_grad():
        gat.bias.normal_()
    gsd = {f"knn_conv.{k}": v.detach().clone() for k, v in gat.state_dict().items()}
    ei = R.knn_to_edge_index(idx)
    y = gat(x, ei)
    assert torch.equal(y, R.gat_conv(gsd, x, ei))
    torch.save({"x": x, "frame_ptr": ptr, "k": 20, "idx": idx, "gat_state_dict": gsd, "gat_out": y.detach()},
               os.path.join(GOLD, "knn_gat_small.pt"))
    print("knn_gat_small ok")


if __name__ == "__main__":
    main()
