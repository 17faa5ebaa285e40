#!/usr/bin/env python
"""Benchmark of the Batch3DMOT tracking-graph GNN hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Metric (BASELINE.json): GNN edges/s for forward + backward. A step = one training pass
(multimodal cl_config model: forward, class-balanced BCE, backward, gradient all-reduce when
N > 1, Adam) over one batch of synthetic nuScenes-shaped scene graphs (configs[1] model on
configs[0]-shaped graphs: T=40, N=2000, E~60k per scene; `--scenes` graphs per rank, default 64). One edge =
one directed input edge taken through the whole model once.

Under torchrun every rank holds its own scenes (weak scaling); time = max over ranks.
Beside the headline the line carries (extra keys): the forward-only figure, the end-to-end figure, the
1e-4-tolerance (fp32) mode, the poses-only model (configs[0]), the k-NN / attention-conv stress (configs[2]),
batched scene-sharded inference + track assembly over 150 val-shaped scenes (configs[3], strong scaling over
the ranks), the reference's own 2-window training batch (configs[4] small-batch regime) and the bf16-vs-fp32
output error of the timed batch.

`--impl reference` times the reference's own CPU implementation on the host cores (rank 0 only): the
UNMODIFIED reference model files from oracle/_ref (written by oracle/make_ref.py in the build container)
under the PyG stand-ins, forward + BCELoss(weight) + backward + torch.optim.Adam as in train.py:126-160;
if oracle/_ref is absent, the restated port (oracle/ref_restated.py) with the same optimiser."""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time
from types import SimpleNamespace

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MM_FWD_FLOPS_PER_EDGE = 4_581_776          # SURVEY §8d (edge terms, depth 6)
MM_FWDBWD_FLOPS_PER_EDGE = 3 * MM_FWD_FLOPS_PER_EDGE
SEED = 5621


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scenes", type=int, default=64, help="scene graphs per rank per step")
    ap.add_argument("--precision", default=os.environ.get("B3D_PRECISION", "bf16"), choices=["bf16", "fp32"],
                    help="bf16: tcgen05 tiles (2e-2 parity mode, north_star); fp32: 1e-4 parity mode")
    ap.add_argument("--cpu-scenes", type=int, default=4, help="scene graphs in the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="headline + e2e + roofline only")
    ap.add_argument("--val-scenes", type=int, default=150, help="configs[3]: scenes in the batched-inference run")
    ap.add_argument("--profile-only", action="store_true", help="timed loop only (for ncu launch lists)")
    return ap.parse_args()


def make_batch(rank, scenes, raw=False, ids=None):
    """`scenes` scene graphs of rank `rank` (seeds SEED + 1000 rank + i), or the scenes `ids` of rank 0's batch."""
    from batch3dmot_b200 import synth
    gs = []
    for i in (range(scenes) if ids is None else ids):
        s = SEED + 1000 * rank + i
        gs.append(synth.add_labels(synth.add_modalities(synth.scene_graph(seed=s), s, raw=raw), s))
    return synth.collate(gs)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ----------------------------------------------------------------------------- reference arm (CPU)
def cpu_reference_run(steps, warmup, n_scenes):
    """The reference's own CPU training step on the host cores: (edges/s, ms/step, E, threads, kind, how)."""
    from batch3dmot_b200 import synth
    from oracle import make_ref
    torch.set_num_threads(os.cpu_count())
    data = make_batch(0, n_scenes, raw=True)
    E = data.edge_index.size(1)
    root = make_ref.ref_root()
    if root is not None:
        from oracle import pyg_shim
        _, clr = pyg_shim.load_reference(root)
        torch.manual_seed(SEED)
        enc = [synth.EmbeddingEncoder(t) for t in (data.x_img, data.pointnet_out, data.radarnet_out)]
        model = clr.GNN(*enc)
        model.train()
        opt = torch.optim.Adam(model.parameters(), lr=1e-4, weight_decay=1e-4, betas=(0.9, 0.999))   # train.py:106-109
        gt = data.y.float()

        def step():
            out, _ = model.forward(data)                                                            # train.py:133
            loss = torch.nn.BCELoss(weight=data.edge_weights)(out.squeeze(1), gt) / 2               # :139-141
            opt.zero_grad()
            loss.backward()                                                                          # :157-160
            opt.step()
        kind = "reference"
        how = ("UNMODIFIED reference clr_att_gnn.py (oracle/_ref) under the PyG stand-ins of oracle/pyg_shim.py "
               "(scatter / knn_graph / GATConv / MessagePassing restated in torch; PyG's C++ ops are not installable), "
               "random-embedding encoder stubs, fwd + BCELoss(weight) + bwd + torch.optim.Adam")
    else:
        from oracle import ref_restated as R
        from batch3dmot_b200.clr_att_gnn import GNN
        torch.manual_seed(SEED)
        sd = GNN(None, None, None).state_dict()
        params = {k: v.clone().requires_grad_(not k.startswith("knn_conv")) for k, v in sd.items()}
        mha = R.build_mha(params)
        opt = torch.optim.Adam([p for p in params.values() if p.requires_grad], lr=1e-4, weight_decay=1e-4)

        def step():
            R.cpu_train_step(params, data, mha)
            opt.step()
        kind = "port"
        how = "oracle/ref_restated.py faithful mode (oracle/_ref absent) + torch.optim.Adam"
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return E / dt, dt * 1e3, E, torch.get_num_threads(), kind, how


# ----------------------------------------------------------------------------- helpers for the GPU arm
class Timer:
    def __init__(self, dev, world):
        self.dev, self.world = dev, world
        self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def barrier(self):
        import torch.distributed as dist
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def run(self, fn, steps, warmup, reduce=True):
        """ms per call of fn(): W untimed calls, then K calls between barriers, CUDA events, max over ranks."""
        import torch.distributed as dist
        for _ in range(warmup):
            fn()
        self.barrier() if reduce else torch.cuda.synchronize()
        self.e0.record()
        for _ in range(steps):
            fn()
        self.e1.record()
        self.barrier() if reduce else torch.cuda.synchronize()
        ms = torch.tensor([self.e0.elapsed_time(self.e1)], device=self.dev)
        if reduce and self.world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / steps


def to_dev(ns, dev):
    return SimpleNamespace(**{k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in vars(ns).items()})


def mm_kwargs(d):
    return dict(x_img=d.x_img, pointnet_out=d.pointnet_out, radarnet_out=d.radarnet_out, lidar_mask=d.m_lidar,
                radar_mask=d.m_radar)


def val_scene_rates(n, seed=SEED):
    """configs[3]: per-scene node rate ~ LogNormal(mean 75 / frame, sigma 0.5) clipped to [5, 300]."""
    g = torch.Generator().manual_seed(seed + 77)
    r = torch.exp(torch.randn(n, generator=g) * 0.5 + math.log(75.0) - 0.125)
    return r.clamp(5, 300).round().long().tolist()


def extras(a, rank, world, dev, model, d, tm, E_global, trainer, headline):
    """The other BASELINE configs and regimes, as extra keys of the JSON line."""
    import torch.distributed as dist
    from batch3dmot_b200 import _lib, ops, synth, inference
    from batch3dmot_b200.pose_gnn import PoseGNN
    from batch3dmot_b200.clr_att_gnn import GNN
    from batch3dmot_b200.parallel import Trainer, lpt_partition
    from batch3dmot_b200.gat import GATConv
    x = {}
    kw = mm_kwargs(d)

    # ---- configs[3]: batched inference over val-shaped scenes, sharded by scene over the ranks (strong scaling):
    # every rank generates and owns its LPT share, cuts windows on the device, one forward per chunk, track assembly
    rates = val_scene_rates(a.val_scenes)
    est = [40 * r * min(40.0, 1.1 * r) for r in rates]          # edge-count estimate known to every rank without generating
    mine = lpt_partition(est, world)[rank]
    scenes = {i: inference.pin_scene(synth.add_modalities(
        synth.scene_graph(seed=SEED + 5000 + i, T=40, frame_sizes=[rates[i]] * 40), SEED + 5000 + i, raw=False)) for i in mine}
    lst = [scenes[i] for i in mine]
    model.eval()
    n_tracks = [0]

    def infer():
        res = inference.track_scenes(model, lst, dev, multimodal=True, want_tracks=False)
        n_tracks[0] = sum(int(v[0].max()) + 1 for v in res.values() if v[0].numel())
    infer()                                                     # warm-up (allocator, weight packs)
    tm.barrier()
    t0 = time.perf_counter()
    infer()
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    win_edges = torch.tensor([sum(int(2.4 * s.edge_index.size(1)) for s in lst)], device=dev)   # ~2.4 windows per edge
    if world > 1:
        dist.all_reduce(win_edges)
    x["batched_inference"] = {
        "workload": f"configs[3]: {a.val_scenes} val-shaped scenes x 36 sliding 5-frame windows, scene-sharded (LPT) over "
                    f"{world} GPU(s): host->device copy of the (pinned) scenes, window cut on device, forward, track assembly",
        "scenes_per_s": a.val_scenes / float(dt.item()), "seconds": float(dt.item()), "scaling": "strong",
        "window_edges_approx": int(win_edges.item()), "tracks_rank0": n_tracks[0], "timed": "host wall clock, max over ranks"}
    model.train()
    del scenes, lst

    # ---- configs[4] as STRONG scaling: the SAME global batch (rank 0's `--scenes` scene graphs of the headline) split
    # over the ranks by scene (contiguous shares), one gradient all-reduce per step; at 1 GPU this is the headline itself
    if world == 1:
        x["strong_scaling_training"] = dict(headline, global_scenes=a.scenes, scenes_per_gpu=a.scenes, note="= the headline step")
    else:
        per = -(-a.scenes // world)
        ids = list(range(rank * per, min(a.scenes, (rank + 1) * per)))
        if ids:
            ds = to_dev(make_batch(0, 0, ids=ids), dev)
            ds._b3d_graph = ops.Graph(ds.edge_index, ds.num_nodes)
            Es = ds.edge_index.size(1)
        else:
            ds, Es = None, 0
        et = torch.tensor([Es], dtype=torch.int64, device=dev)
        dist.all_reduce(et)
        Eg = int(et.item())
        if (world - 1) * per < a.scenes:      # every rank owns at least one scene (same decision on every rank)
            ms_s = tm.run(lambda: trainer.step(ds, global_edges=Eg, **mm_kwargs(ds)), max(3, a.steps // 2), 2)
            x["strong_scaling_training"] = {"global_scenes": a.scenes, "scenes_per_gpu": per, "edges_per_step": Eg,
                                            "ms_per_step": ms_s, "edges_per_s": Eg / ms_s * 1e3, "scaling": "strong",
                                            "timed": "CUDA events, max over ranks"}
            # the same step as a CUDA graph per rank (Trainer.capture: kernels + the NCCL all-reduce + Adam captured): with
            # few scenes per GPU the eager step is bound by the host-side dispatch of its ~490 launches, not by the GPU
            if os.environ.get("B3D_BENCH_GRAPH", "1") == "1":
                ok = torch.ones(1, device=dev)
                try:
                    replay = trainer.capture(ds, global_edges=Eg, **mm_kwargs(ds))
                except Exception as ex:                      # noqa: BLE001 - reported, not fatal for the headline
                    ok.zero_()
                    x["strong_scaling_training"]["cuda_graph_error"] = repr(ex)[:200]
                dist.all_reduce(ok, op=dist.ReduceOp.MIN)    # replay only when every rank holds a graph (collective inside)
                if float(ok.item()) > 0:
                    ms_g = tm.run(replay, max(3, a.steps // 2), 2)
                    x["strong_scaling_training"].update(cuda_graph_ms_per_step=ms_g, cuda_graph_edges_per_s=Eg / ms_g * 1e3)
                    del replay
                trainer._step_dev = None
        del ds
        # the reference's OWN batch on every rank (cl_config.yaml:99: 2 window graphs), data parallel: ~490 launches of a
        # few microseconds + one 5.28 MB all-reduce per step; eager (host-dispatch bound) and as a CUDA graph per rank
        sd = SEED + 31 * rank
        wins = synth.windows(synth.add_labels(synth.add_modalities(synth.scene_graph(seed=sd), sd, raw=False), sd), 5)[:2]
        for w in wins:
            synth.add_labels(w, sd)
        dw = to_dev(synth.collate(wins), dev)
        dw._b3d_graph = ops.Graph(dw.edge_index, dw.num_nodes)
        et = torch.tensor([dw.edge_index.size(1)], dtype=torch.int64, device=dev)
        dist.all_reduce(et)
        Ew = int(et.item())
        kww = mm_kwargs(dw)
        ms_e = tm.run(lambda: trainer.step(dw, global_edges=Ew, **kww), 10, 3)
        res = {"windows_per_gpu": 2, "edges_per_step": Ew, "us_per_step": ms_e * 1e3, "edges_per_s": Ew / ms_e * 1e3,
               "scaling": "weak"}
        if os.environ.get("B3D_BENCH_GRAPH", "1") == "1":
            ok = torch.ones(1, device=dev)
            try:
                replay = trainer.capture(dw, global_edges=Ew, **kww)
            except Exception as ex:                          # noqa: BLE001
                ok.zero_()
                res["cuda_graph_error"] = repr(ex)[:200]
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if float(ok.item()) > 0:
                ms_g = tm.run(replay, 20, 3)
                res.update(cuda_graph_us_per_step=ms_g * 1e3, cuda_graph_edges_per_s=Ew / ms_g * 1e3)
                del replay
            trainer._step_dev = None
        x["small_batch_data_parallel"] = res
        del dw
    if rank != 0:
        return x

    # ---- bf16 output error of the timed batch against the fp32 (1e-4) mode, once
    if a.precision == "bf16":
        with torch.no_grad():
            out16 = model(d, **kw)[0].float()
            ops.set_precision("fp32")
            out32 = model(d, **kw)[0].float()
            ops.set_precision("bf16")
        diff = (out16 - out32).abs()
        x["parity_timed_batch"] = {"max_abs_err": float(diff.max()), "max_rel_to_max": float(diff.max() / out32.abs().max()),
                                   "mean_rel_elementwise": float((diff / out32.abs().clamp_min(1e-6)).mean()),
                                   "against": "fp32 mode of the same kernels' library on the same batch (inference forward)"}
        del out16, out32, diff

    # ---- the 1e-4-tolerance mode (fp32 parity arithmetic) on 8 scenes
    small = to_dev(make_batch(0, 8), dev)
    small._b3d_graph = ops.Graph(small.edge_index, small.num_nodes)
    Es = small.edge_index.size(1)
    ops.set_precision("fp32")
    torch.manual_seed(SEED)
    m32 = GNN(None, None, None).to(dev)
    tr32 = Trainer(m32, batch_size=2, data_parallel=False)        # rank-0-only extras: no collectives
    ms = tm.run(lambda: tr32.step(small, **mm_kwargs(small)), 3, 2, reduce=False)
    with torch.no_grad():
        msf = tm.run(lambda: m32(small, **mm_kwargs(small)), 3, 1, reduce=False)
    x["fp32_mode"] = {"fwd_bwd_edges_per_s": Es / ms * 1e3, "forward_edges_per_s": Es / msf * 1e3, "scenes": 8, "edges": Es,
                      "tolerance": "1e-4 relative (tests/test_gpu_models.py)"}
    ops.set_precision(a.precision)
    del m32, tr32

    # ---- configs[0]: poses-only model, forward and forward+backward, 64 scenes and 1 scene
    torch.manual_seed(SEED)
    pm = PoseGNN().to(dev)
    ptr_ = Trainer(pm, batch_size=2, from_logits=True, data_parallel=False)
    one = to_dev(synth.add_labels(synth.scene_graph(seed=SEED), SEED), dev)
    one._b3d_graph = ops.Graph(one.edge_index, one.num_nodes)
    res = {}
    for name, dd in (("64_scenes", d), ("1_scene", one)):
        Ed = dd.edge_index.size(1)
        ms = tm.run(lambda: ptr_.step(dd), 5, 3, reduce=False)
        with torch.no_grad():
            msf = tm.run(lambda: pm(dd), 5, 2, reduce=False)
        res[name] = {"edges": Ed, "fwd_bwd_edges_per_s": Ed / ms * 1e3, "forward_edges_per_s": Ed / msf * 1e3,
                     "ms_per_step": ms}
        if name == "1_scene":        # configs[0] proper: one scene graph per call is launch-bound -> CUDA-graph replay
            try:
                rp = inference.capture_forward(pm, dd, multimodal=False)
                msg = tm.run(rp, 20, 3, reduce=False)
                res[name].update(cuda_graph_forward_us=msg * 1e3, cuda_graph_forward_edges_per_s=Ed / msg * 1e3)
                del rp
            except Exception as ex:                  # noqa: BLE001 - reported, not fatal for the headline
                res[name]["cuda_graph_error"] = repr(ex)[:200]
    x["pose_model"] = dict(res, workload="configs[0]: poses-only PoseGNN, random init, fwd + BCE-with-logits + bwd + Adam")
    del pm, ptr_

    # ---- configs[4] small-batch regime: the reference's own training batch (cl_config.yaml:99 batch_size 2 =
    # two 5-frame window graphs) and one scene graph
    wins = synth.windows(synth.add_labels(synth.add_modalities(synth.scene_graph(seed=SEED), SEED, raw=False), SEED), 5)[:2]
    for w in wins:
        synth.add_labels(w, SEED)
    two = to_dev(synth.collate(wins), dev)
    one_mm = to_dev(make_batch(0, 1), dev)
    all_w = synth.windows(synth.add_labels(synth.add_modalities(synth.scene_graph(seed=SEED), SEED, raw=False), SEED), 5)
    for w in all_w:
        synth.add_labels(w, SEED)
    scene_w = to_dev(synth.collate(all_w), dev)          # SURVEY 8d config 5 large-batch variant: all windows of one scene
    torch.manual_seed(SEED)
    ms_ = GNN(None, None, None).to(dev)
    trs = Trainer(ms_, batch_size=2, data_parallel=False)
    res = {}
    for name, dd in (("2_windows", two), ("1_scene", one_mm), ("all_windows_of_1_scene", scene_w)):
        dd._b3d_graph = ops.Graph(dd.edge_index, dd.num_nodes)
        Ed = dd.edge_index.size(1)
        kwd = mm_kwargs(dd)
        trs.step(dd, **kwd)
        _lib.reset_launch_count()
        trs.step(dd, **kwd)
        lc = _lib.launch_count()
        ms = tm.run(lambda: trs.step(dd, **kwd), 10, 3, reduce=False)
        res[name] = {"edges": Ed, "us_per_step": ms * 1e3, "fwd_bwd_edges_per_s": Ed / ms * 1e3, "libb3d_launches_per_step": lc}
        try:                                         # the same step captured in a CUDA graph (Trainer.capture)
            replay = trs.capture(dd, **kwd)
            msg = tm.run(replay, 20, 3, reduce=False)
            res[name].update(cuda_graph_us_per_step=msg * 1e3, cuda_graph_edges_per_s=Ed / msg * 1e3)
            del replay
        except Exception as ex:                      # noqa: BLE001 - reported, not fatal for the headline
            res[name]["cuda_graph_error"] = repr(ex)[:200]
        trs._step_dev = None
    # a STREAM of different 2-window batches (what train.py's DataLoader delivers): eager Trainer.step against
    # parallel.BucketedTrainer (batches padded to size buckets, one captured CUDA graph per bucket, replayed)
    try:
        from batch3dmot_b200.parallel import BucketedTrainer
        stream = []
        for sd in (SEED + 1, SEED + 2):
            ws = synth.windows(synth.add_labels(synth.add_modalities(synth.scene_graph(seed=sd), sd, raw=False), sd), 5)
            for w in ws:
                synth.add_labels(w, sd)
            stream += [to_dev(synth.collate(ws[i:i + 2]), dev) for i in range(0, 24, 2)]
        e_stream = sum(b.edge_index.size(1) for b in stream)

        def run_stream(step):
            for b in stream:
                step(b, **mm_kwargs(b))
        ms_eager = tm.run(lambda: run_stream(trs.step), 1, 1, reduce=False)
        bt = BucketedTrainer(trs)
        run_stream(bt.step)                                  # first pass: captures one graph per size bucket
        ms_b = tm.run(lambda: run_stream(bt.step), 2, 1, reduce=False)
        res["stream_of_2_window_batches"] = {
            "batches": len(stream), "edges": e_stream, "size_buckets": bt.captures,
            "eager_us_per_step": ms_eager * 1e3 / len(stream), "bucketed_graph_us_per_step": ms_b * 1e3 / len(stream),
            "eager_edges_per_s": e_stream / ms_eager * 1e3, "bucketed_graph_edges_per_s": e_stream / ms_b * 1e3}
        del bt, stream
    except Exception as ex:                                  # noqa: BLE001 - reported, not fatal for the headline
        res["stream_of_2_window_batches"] = {"error": repr(ex)[:200]}
    trs._step_dev = None
    x["small_batch"] = dict(res, workload="configs[4]: multimodal training step at the reference's batch (2 window graphs) and at 1 scene")
    del ms_, trs

    # ---- configs[2]: frame-wise k-NN + attention-weighted conv stress (N = 200k nodes, frames of 250)
    res = {}
    for D in (48, 96):
        xk, ptr = synth.knn_stress(SEED, 200_000, frame=250, D=D)
        xk, ptr = xk.to(dev), ptr.to(dev)
        conv = GATConv(D, D).to(dev)
        for k in (8, 16):
            nbr = ops.knn_frames(xk, ptr, k)
            us_knn = tm.run(lambda: ops.knn_frames(xk, ptr, k), 5, 2, reduce=False) * 1e3
            with torch.no_grad():
                us_gat = tm.run(lambda: conv.forward_table(xk, nbr), 5, 2, reduce=False) * 1e3
            nbytes = 200_000 * (D * 4 + k * 8)                     # read x once, write the neighbour table
            res[f"D{D}_k{k}"] = {"knn_us": us_knn, "knn_algorithmic_gbs": nbytes / us_knn / 1e3,
                                 "knn_gflops": 200_000 * 250 * D * 3 / us_knn / 1e3, "gat_us": us_gat}
    x["knn_attention_conv"] = dict(res, workload="configs[2]: 200k nodes in frames of 250; brute-force frame-local k-NN "
                                                 "(sub, mul, add per feature) + GATConv lin / softmax / aggregate")
    return x


def main():
    a = parse()
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    config = {"workload": f"configs[1] multimodal cl_config GNN training step (fwd + cb-BCE + bwd + Adam) on "
                          f"{a.scenes} synthetic nuScenes-shaped scene graphs per GPU (T=40, N=2000, E~60k each)",
              "scenes_per_gpu": a.scenes, "precision": a.precision, "parallelism": f"scene-sharded dp{world}",
              "l2": "inputs + saved activations per step exceed the 126 MB L2"}

    if a.impl == "reference":
        if rank != 0:
            return
        steps, warm = max(1, min(a.steps, 3)), max(1, min(a.warmup, 1))
        v, ms, E, thr, kind, how = cpu_reference_run(steps, warm, a.cpu_scenes)
        print(json.dumps({
            "impl": "reference", "metric": "gnn_edges_per_s_fwd_bwd", "value": v, "unit": "edges/s", "n_gpus": a.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(config, note=f"bounded sample: {a.cpu_scenes} of the {a.scenes} scene graphs per step; edges/s "
                                        "normalises the batch size"),
            "cpu_baseline": {"value": v, "unit": "edges/s", "cores": thr, "kind": kind,
                             "sample": f"{a.cpu_scenes} scene graphs, {E} edges per step, {steps} steps; {how}"},
            "e2e": {"value": v, "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch.distributed as dist
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=300))
    from batch3dmot_b200 import _lib, ops, build
    from batch3dmot_b200.clr_att_gnn import GNN
    from batch3dmot_b200.parallel import Trainer
    if local == 0:
        build.build()          # no-op when the in-tree libb3d.so is up to date (one rank per node builds)
    if world > 1:
        dist.barrier()
    _lib.lib()   # fail loudly if the CUDA library is missing
    tm = Timer(dev, world)

    host = make_batch(rank, a.scenes)
    E, N = host.edge_index.size(1), host.num_nodes
    keys = ["pose_feats", "edge_index", "edge_attr", "x_img", "pointnet_out", "radarnet_out", "m_lidar", "m_radar",
            "y", "edge_weights", "node_timestamps"]
    pinned = {k: getattr(host, k).pin_memory() for k in keys}
    h2d_bytes = sum(t.numel() * t.element_size() for t in pinned.values())

    def to_device():
        return SimpleNamespace(**{k: t.to(dev, non_blocking=True) for k, t in pinned.items()}, num_nodes=N)

    e_tot = torch.tensor([E], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(e_tot)
    E_global = int(e_tot.item())

    torch.manual_seed(SEED)
    model = GNN(None, None, None).to(dev)
    ops.set_precision(a.precision)
    trainer = Trainer(model, batch_size=2)

    # ---- device-resident timing ("value")
    d = to_device()
    d._b3d_graph = ops.Graph(d.edge_index, N)
    kw = mm_kwargs(d)
    for _ in range(a.warmup):
        trainer.step(d, global_edges=E_global, **kw)
    tm.barrier()
    _lib.reset_launch_count()
    ev0, ev1 = tm.e0, tm.e1
    with ClockSampler(local) as clk:
        ev0.record()
        for _ in range(a.steps):
            loss = trainer.step(d, global_edges=E_global, **kw)
        ev1.record()
        tm.barrier()
    launches = _lib.launch_count()
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms.item()) / a.steps
    value = E_global / (ms_per_step * 1e-3)

    if a.profile_only:
        if rank == 0:
            print(json.dumps({"value": value, "ms_per_step": ms_per_step, "gpu_launches": launches}))
        return
    # ---- forward only (inference, configs[3]'s per-rank work): same batch, no autograd, nothing saved
    model.eval()
    with torch.no_grad():
        msf = tm.run(lambda: model(d, **kw), a.steps, 2)
    model.train()
    fwd_value = E_global / (msf * 1e-3)

    # ---- end-to-end through the public API with host buffers ("e2e"): every step copies ITS inputs
    # from pinned host memory (on a copy stream, overlapped with the previous step's kernels), builds
    # the CSR from the fresh edge_index, runs the step and reads the loss back to the host.
    # Two preallocated device staging sets (double buffer): no allocation inside the loop; a set is
    # overwritten only after the step that read it has finished (event), and the in-place copy bumps the
    # tensor version, so the CSR cache misses and the tables are rebuilt from the fresh edge_index.
    copy_stream = torch.cuda.Stream(device=dev)
    bufs = [{k: torch.empty(t.shape, dtype=t.dtype, device=dev) for k, t in pinned.items()} for _ in range(2)]
    done = [None, None]

    def stage_inputs(slot):
        with torch.cuda.stream(copy_stream):
            if done[slot] is not None:
                copy_stream.wait_event(done[slot])
            for k, t in pinned.items():
                bufs[slot][k].copy_(t, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(copy_stream)
        return SimpleNamespace(**bufs[slot], num_nodes=N), ready

    def e2e_loop(n):
        nxt = stage_inputs(0)
        last = None
        for i in range(n):
            dd, ready = nxt
            torch.cuda.current_stream().wait_event(ready)
            if i + 1 < n:
                nxt = stage_inputs((i + 1) & 1)             # H2D of step i+1 overlaps step i
            l = trainer.step(dd, global_edges=E_global, **mm_kwargs(dd))   # builds CSR from edge_index
            done[i & 1] = torch.cuda.Event()
            done[i & 1].record()
            last = float(l.item())                          # D2H of the loss (host sync every step)
        return last

    e2e_loop(max(3, a.warmup))     # warm-up: both staging buffer sets and the per-step CSR tables get allocated
    tm.barrier()
    ev0.record()
    e2e_loop(a.steps)
    ev1.record()
    tm.barrier()
    ms2 = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_value = E_global / (float(ms2.item()) / a.steps * 1e-3)
    del bufs

    # ---- roofline of the dominant kernel
    hbm, tf_burst, tf_sust, src = peaks()

    def time_kernel(fn, reps=10):
        return tm.run(fn, reps, 3, reduce=False) * 1e-3

    if a.precision == "bf16":
        # dominant kernel class: k_linear_tma (TMA-fed tcgen05 tiles); heaviest launch = att_edge_encoder layer 2
        # ([E,512] bf16 -> 384, ReLU, bf16 out). Algorithmic FLOPs = 2*E*512*384; algorithmic bytes = E*(512+384)*2.
        # Timed alone on a fixed 489,812-row operand (the shape of the committed ncu capture).
        lin = model.att_edge_encoder[2]
        Er = 489812
        h = torch.randn(Er, 512, device=dev).to(torch.bfloat16)
        out = torch.empty(Er, 384, device=dev, dtype=torch.bfloat16)
        t = time_kernel(lambda: ops.linear_raw([(h, None, None, 0)], lin.weight, lin.bias, Er, 1, out=out, tc=True))
        # Intensity 2*512*384 / ((512+384)*2) = 219 FLOP/B is below the measured ridge (1635 TFLOP/s / 6.55 TB/s
        # = 250 FLOP/B): the binding roof of this streaming GEMM is HBM; the tensor-pipe view is kept beside it.
        ach_t = 2.0 * Er * 512 * 384 / t / 1e12
        ach = Er * (512 + 384) * 2 / t / 1e9
        roof = {"kernel": "k_linear_tma<relu> (TMA-fed tcgen05 bf16; att_edge_encoder layer 2, [E,512]x[512,384])", "bound": "hbm",
                "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
                "traffic": 849_700_000,   # dram read + write bytes per launch (502.1 MB + 347.6 MB), ncu --set full of the lean build, same shape
                "traffic_source": "profiles/r2_roofline_kernel.md, r2_lean_epilogue.md last table row (algorithmic bytes: E*(512+384)*2 = 877,743,104)",
                "peak_source": src + " HBM copy bandwidth (burst); kernel timed alone with CUDA events",
                "rows": Er, "us_per_launch": t * 1e6,
                "tensor_view": {"achieved_tflops": ach_t, "peak_tflops": tf_burst, "frac": ach_t / tf_burst}}
        del h, out
    else:
        # fp32 parity path: the dominant launch is the edge_update layer 0 ([E,320] -> 256, gathered)
        mp = model.message_passing
        lin = mp.edge_update[0]
        x = torch.randn(N, 96, device=dev); e = torch.randn(E, 64, device=dev); att = torch.randn(E, 64, device=dev)
        g = d._b3d_graph
        items = [(x, g.dst32, None, 0), (x, g.src32, None, 0), (e, None, None, 0), (att, None, None, 0)]
        out = torch.empty(E, 256, device=dev)
        t = time_kernel(lambda: ops.linear_raw(items, lin.weight, lin.bias, E, 1, out=out))
        ach = 2.0 * E * 320 * 256 / t / 1e12
        roof = {"kernel": "edge_update layer 0, gathered [E,320]x[320,256] (fp32 parity mode)", "bound": "tensor",
                "achieved": ach, "peak": tf_burst, "unit": "TFLOP/s", "frac": ach / tf_burst, "traffic": None,
                "peak_source": src + " bf16 dense burst; kernel timed alone"}
        del x, e, att, out

    # ---- step-level view of the same roofline: every edge-level dense-layer launch of ONE extra step timed with
    # CUDA events around the launch (ops.optime: synchronises per launch, so this step is not part of any
    # throughput figure); algorithmic operand bytes (inputs + outputs + masks + addends) / summed time
    roof_step = None
    if a.precision == "bf16":
        ops.optime_begin()
        trainer.step(d, global_edges=E_global, **kw)
        rec = ops.optime_end()
        agg = {}
        for sig, (n, ms_k, nb) in rec.items():
            if sig[1] != E:            # edge-level launches only (node-level ones are latency-bound and small)
                continue
            cls = "k_wgrad_tma" if sig[0] == "wgrad" else "k_linear_tma"
            ent = agg.setdefault(cls, [0, 0.0, 0, 0])
            # row-gathered addends ([E, n_out] bf16 rows read from small per-node tables) are served by L2, not by HBM:
            # they are counted in the operand bytes and left out of the HBM-only view
            gathered = n * sig[7] * sig[1] * sig[3] * 2
            ent[0] += n; ent[1] += ms_k; ent[2] += nb; ent[3] += nb - gathered
        roof_step = {cls: {"launches": n, "ms": ms_k, "algorithmic_gb": nb / 1e9, "achieved_gbs": nb / ms_k / 1e6,
                           "frac_of_hbm_peak": nb / ms_k / 1e6 / hbm,
                           "hbm_only": {"algorithmic_gb": nh / 1e9, "achieved_gbs": nh / ms_k / 1e6,
                                        "frac_of_hbm_peak": nh / ms_k / 1e6 / hbm,
                                        "note": "without the row-gathered addend bytes (L2-resident node tables)"}}
                     for cls, (n, ms_k, nb, nh) in agg.items() if ms_k > 0}

    line = {"metric": "gnn_edges_per_s_fwd_bwd", "value": value, "unit": "edges/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if a.precision == "fp32" else "bf16", "data": "synthetic",
            "config": dict(config, edges_per_step=E_global, nodes_per_gpu=N),
            "e2e": {"value": e2e_value, "unit": "edges/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4},
            "gpu_launches": launches, "clocks": clk.summary(), "roofline": roof, "roofline_step": roof_step,
            "algorithmic_tflops": value * MM_FWDBWD_FLOPS_PER_EDGE / 1e12, "loss": float(loss.item()),
            "forward_only": {"value": fwd_value, "unit": "edges/s", "ms_per_step": msf,
                             "note": "inference forward of the same batch under torch.no_grad()"},
            "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 2**30}
    if not a.no_extras:
        line.update(extras(a, rank, world, dev, model, d, tm, E_global, trainer,
                           {"edges_per_step": E_global, "ms_per_step": ms_per_step, "edges_per_s": value}))
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        v, cms, cE, thr, kind, how = cpu_reference_run(1, 1, a.cpu_scenes)
        line["cpu_baseline"] = {"value": v, "unit": "edges/s", "cores": thr, "kind": kind,
                                "sample": f"{a.cpu_scenes} scene graphs, {cE} edges per step, 1 warm-up + 1 timed step "
                                          f"({cms:.0f} ms/step); {how}"}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
