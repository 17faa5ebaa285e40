"""Where does the end-to-end step (pinned host buffers -> loss on host) spend its extra time?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from types import SimpleNamespace
from batch3dmot_b200 import ops
from batch3dmot_b200.clr_att_gnn import GNN
from batch3dmot_b200.parallel import Trainer
dev = torch.device("cuda", 0)
ops.set_precision("bf16")
scenes = int(sys.argv[1]) if len(sys.argv) > 1 else 32
host = bench.make_batch(0, scenes)
N = host.num_nodes
keys = ["pose_feats", "edge_index", "edge_attr", "x_img", "pointnet_out", "radarnet_out", "m_lidar", "m_radar", "y", "edge_weights", "node_timestamps"]
pinned = {k: getattr(host, k).pin_memory() for k in keys}
def to_device():
    return SimpleNamespace(**{k: t.to(dev, non_blocking=True) for k, t in pinned.items()}, num_nodes=N)
torch.manual_seed(0)
tr = Trainer(GNN(None, None, None).to(dev), batch_size=2)
kw = lambda d: dict(x_img=d.x_img, pointnet_out=d.pointnet_out, radarnet_out=d.radarnet_out, lidar_mask=d.m_lidar, radar_mask=d.m_radar)
d = to_device(); d._b3d_graph = ops.Graph(d.edge_index, N)
for _ in range(3): tr.step(d, **kw(d))
torch.cuda.synchronize()
def timeit(name, fn, n=5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); print(f"{name}: {(time.perf_counter() - t0) / n * 1e3:.2f} ms")
timeit("step, cached graph, no sync", lambda: tr.step(d, **kw(d)))
timeit("step, cached graph, loss.item()", lambda: float(tr.step(d, **kw(d)).item()))
timeit("H2D of all inputs (pinned, default stream)", lambda: to_device())
timeit("CSR build (ops.Graph)", lambda: ops.Graph(d.edge_index, N))
def fresh():
    dd = to_device()
    return float(tr.step(dd, **kw(dd)).item())
timeit("H2D + CSR + step + loss.item() (serial)", fresh)
t0 = time.perf_counter(); tr.step(d, **kw(d)); t1 = time.perf_counter(); torch.cuda.synchronize()
print(f"host time to enqueue one step: {(t1 - t0) * 1e3:.2f} ms")

# ---- overlapped staging exactly as bench.py does it
copy_stream = torch.cuda.Stream(device=dev)
def stage():
    with torch.cuda.stream(copy_stream):
        dd = to_device()
        ready = torch.cuda.Event(); ready.record(copy_stream)
    return dd, ready
def loop(n, rec=True):
    nxt = stage()
    for i in range(n):
        dd, ready = nxt
        torch.cuda.current_stream().wait_event(ready)
        if rec:
            for t in vars(dd).values():
                if torch.is_tensor(t): t.record_stream(torch.cuda.current_stream())
        if i + 1 < n: nxt = stage()
        float(tr.step(dd, **kw(dd)).item())
loop(2)
timeit("overlapped loop (bench e2e), per step", lambda: loop(6), n=1)
print("  ^ divide by 6")
# H2D concurrently with a step: how long does each take?
e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
torch.cuda.synchronize()
e[0].record(); 
with torch.cuda.stream(copy_stream):
    e[2].record(copy_stream); dd = to_device(); e[3].record(copy_stream)
tr.step(d, **kw(d)); e[1].record(); torch.cuda.synchronize()
print(f"concurrent: step {e[0].elapsed_time(e[1]):.2f} ms, H2D {e[2].elapsed_time(e[3]):.2f} ms")
