"""GNN.forward(data) fed by the reference's REAL modality encoders — ResNetAE.encode,
PointNetClassifier.forward_feat, RadarNetClassifier.forward_feat (models/resnet_fully_conv.py:155-161,
pointnet.py:168-192, radarnet.py:40-64), consumed at clr_att_gnn.py:125-141 — against the unmodified reference GNN
under the PyG stand-ins with the same encoder objects' weights. The encoder classes come from oracle/_ref (the
reference's own files, shipped by oracle/make_ref.py); they are test infrastructure here: the product only
duck-types them."""
import copy
from types import SimpleNamespace

import pytest
import torch

from oracle import make_ref
from batch3dmot_b200 import synth

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(make_ref.ref_root() is None, reason="oracle/_ref not present")]
DEV = "cuda"


def _data(seed=31, T=6, npf=12):
    g = torch.Generator().manual_seed(seed)
    d = synth.scene_graph(seed=seed, T=T, nodes_per_frame=npf, k=6)
    N = d.pose_feats.size(0)
    d.img_feats = torch.randn(N, 3, 32, 32, generator=g)
    d.lidar_feats = torch.randn(N, 128, 3, generator=g) * (torch.rand(N, 1, 1, generator=g) < 0.7)
    d.radar_feats = torch.randn(N, 64, 4, generator=g) * (torch.rand(N, 1, 1, generator=g) < 0.3)
    return d


@pytest.mark.parametrize("seed", [31, 32])
def test_gnn_forward_with_real_encoders_matches_reference(seed):
    from oracle import pyg_shim
    from batch3dmot_b200.clr_att_gnn import GNN
    _, clr = pyg_shim.load_reference(make_ref.ref_root())
    torch.manual_seed(seed)
    encs = [clr.resnet_fully_conv.ResNetAE(), clr.pointnet.PointNetClassifier(), clr.radarnet.RadarNetClassifier()]
    ref = clr.GNN(*encs).eval()                     # eval: BatchNorm running stats / no dropout in the encoders
    data = _data(seed)
    with torch.no_grad():
        out_ref, xs_ref = ref(data)
    assert xs_ref.shape == (data.pose_feats.size(0), 288)
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        m = GNN(*[copy.deepcopy(e) for e in encs]).eval()
        m.load_state_dict(ref.state_dict(), strict=True)   # encoder weights travel under resnet.* / pointnet.* / radarnet.*
        m = m.to(DEV)
        d = SimpleNamespace(**{k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in vars(data).items()})
        with torch.no_grad():
            out, xs = m(d)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    rel = lambda a, b: float((a.double().cpu() - b.double()).abs().max() / b.double().abs().max())
    assert rel(xs, xs_ref) < 1e-4
    assert rel(out, out_ref) < 1e-4
    # missing modalities give exactly-zero embedding blocks (clr_att_gnn.py:131-141), camera is never masked
    miss_l = (data.lidar_feats.reshape(len(xs), -1).sum(1) == 0)
    assert bool((xs.cpu()[miss_l][:, 96:224] == 0).all()) and bool(miss_l.any())
