"""torch.profiler view of one training step (debug): which torch ops (with shapes) own the
non-libb3d elementwise kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from batch3dmot_b200 import ops
from batch3dmot_b200.clr_att_gnn import GNN
from batch3dmot_b200.parallel import Trainer
from types import SimpleNamespace
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda", 0)
ops.set_precision("bf16")
scenes = int(sys.argv[1]) if len(sys.argv) > 1 else 8
host = bench.make_batch(0, scenes)
d = SimpleNamespace(**{k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in vars(host).items()})
d._b3d_graph = ops.Graph(d.edge_index, d.num_nodes)
torch.manual_seed(5621)
tr = Trainer(GNN(None, None, None).to(dev), batch_size=2)
kw = dict(x_img=d.x_img, pointnet_out=d.pointnet_out, radarnet_out=d.radarnet_out, lidar_mask=d.m_lidar, radar_mask=d.m_radar)
for _ in range(2):
    tr.step(d, **kw)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    tr.step(d, **kw)
    torch.cuda.synchronize()
print(prof.key_averages(group_by_input_shape=True).table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=50, max_shapes_column_width=70))
