"""Small-batch regime probe (debug): multimodal training step on 2 window graphs / 1 scene / all windows of a scene,
eager and as a CUDA graph, under the current B3D_FEATURES. Prints us per step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from batch3dmot_b200 import ops, synth, _lib
from batch3dmot_b200.clr_att_gnn import GNN
from batch3dmot_b200.parallel import Trainer
dev = torch.device("cuda", 0)
ops.set_precision("bf16")
SEED = bench.SEED
tm = bench.Timer(dev, 1)
all_w = synth.windows(synth.add_labels(synth.add_modalities(synth.scene_graph(seed=SEED), SEED, raw=False), SEED), 5)
for w in all_w:
    synth.add_labels(w, SEED)
cases = {"2_windows": bench.to_dev(synth.collate(all_w[:2]), dev), "1_scene": bench.to_dev(bench.make_batch(0, 1), dev),
         "all_windows": bench.to_dev(synth.collate(all_w), dev)}
torch.manual_seed(SEED)
tr = Trainer(GNN(None, None, None).to(dev), batch_size=2, data_parallel=False)
print("features", {k: v for k, v in ops.FEATURES.items() if v})
for name, dd in cases.items():
    dd._b3d_graph = ops.Graph(dd.edge_index, dd.num_nodes)
    kw = bench.mm_kwargs(dd)
    tr.step(dd, **kw)
    _lib.reset_launch_count(); tr.step(dd, **kw); lc = _lib.launch_count()
    ms = tm.run(lambda: tr.step(dd, **kw), 10, 3, reduce=False)
    rp = tr.capture(dd, **kw)
    msg = tm.run(rp, 30, 5, reduce=False)
    del rp
    tr._step_dev = None
    print(f"{name:12s} E={dd.edge_index.size(1):7d} launches={lc:4d} eager {ms*1e3:8.1f} us  graph {msg*1e3:8.1f} us")
