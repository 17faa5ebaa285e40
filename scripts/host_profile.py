"""cProfile of the HOST side of eager training steps on the reference's 2-window batch (debug)."""
import cProfile, pstats, os, sys
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
dd = bench.to_dev(synth.collate(all_w[:2]), dev)
dd._b3d_graph = ops.Graph(dd.edge_index, dd.num_nodes)
torch.manual_seed(SEED)
tr = Trainer(GNN(None, None, None).to(dev), batch_size=2, data_parallel=False)
kw = bench.mm_kwargs(dd)
for _ in range(5):
    tr.step(dd, **kw)
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(20):
    tr.step(dd, **kw)
torch.cuda.synchronize()
print("ms per eager step", (time.perf_counter() - t0) / 20 * 1e3)
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    tr.step(dd, **kw)
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(35)
