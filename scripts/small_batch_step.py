"""Three eager multimodal training steps on the reference's 2-window batch (for an ncu launch list)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from batch3dmot_b200 import ops, synth
from batch3dmot_b200.clr_att_gnn import GNN
from batch3dmot_b200.parallel import Trainer
dev = torch.device("cuda", 0)
ops.set_precision("bf16")
SEED = bench.SEED
all_w = synth.windows(synth.add_labels(synth.add_modalities(synth.scene_graph(seed=SEED), SEED, raw=False), SEED), 5)
for w in all_w:
    synth.add_labels(w, SEED)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dd = bench.to_dev(synth.collate(all_w[:n]), dev)
dd._b3d_graph = ops.Graph(dd.edge_index, dd.num_nodes)
torch.manual_seed(SEED)
tr = Trainer(GNN(None, None, None).to(dev), batch_size=2, data_parallel=False)
for _ in range(3):
    tr.step(dd, **bench.mm_kwargs(dd))
torch.cuda.synchronize()
print("E", dd.edge_index.size(1), "N", dd.num_nodes)
