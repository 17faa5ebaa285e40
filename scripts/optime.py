"""Per-op timing of one training step (debug; synchronises after every dense-layer launch).
Prints time, algorithmic bytes and GB/s per (kind, M, K, N, dtypes) signature."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from batch3dmot_b200 import ops
from batch3dmot_b200.clr_att_gnn import GNN
from batch3dmot_b200.parallel import Trainer
from types import SimpleNamespace
dev = torch.device("cuda", 0)
ops.set_precision("bf16")
host = bench.make_batch(0, 8)
d = SimpleNamespace(**{k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in vars(host).items()})
d._b3d_graph = ops.Graph(d.edge_index, d.num_nodes)
torch.manual_seed(5621)
tr = Trainer(GNN(None, None, None).to(dev), batch_size=2)
kw = dict(x_img=d.x_img, pointnet_out=d.pointnet_out, radarnet_out=d.radarnet_out, lidar_mask=d.m_lidar, radar_mask=d.m_radar)
for _ in range(2):
    tr.step(d, **kw)
ops.optime_begin()
tr.step(d, **kw)
r = ops.optime_end()
tot = sum(v[1] for v in r.values())
print(f"dense-layer launches total {tot:.2f} ms")
for sig, (n, ms, nb) in sorted(r.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{ms:7.3f} ms n={n:3d} avg={ms/n*1e3:7.1f} us  {nb/ms/1e6:7.0f} GB/s  {sig}")
