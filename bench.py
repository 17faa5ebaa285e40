#!/usr/bin/env python
"""Benchmark of the Batch3DMOT tracking-graph GNN hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Metric (BASELINE.json): GNN edges/s for forward + backward. A step = one training pass
(multimodal cl_config model: forward, class-balanced BCE, backward, gradient all-reduce when
N > 1, Adam) over one batch of synthetic nuScenes-shaped scene graphs (configs[1] model on
configs[0]-shaped graphs: T=40, N=2000, E~60k per scene; `--scenes` graphs per rank, default 64). One edge =
one directed input edge taken through the whole model once.

Under torchrun every rank holds its own scenes (weak scaling); time = max over ranks.
`--impl reference` times the CPU port of the reference's op stream (oracle/ref_restated.py,
faithful mode) on the host cores — rank 0 only."""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

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
                    help="bf16: tcgen05 tiles (2e-2 parity mode, north_star); fp32: FFMA exact mode (1e-4)")
    ap.add_argument("--cpu-scenes", type=int, default=1, help="scene graphs in the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-only", action="store_true", help="timed loop only (for ncu launch lists)")
    return ap.parse_args()


def make_batch(rank, scenes):
    from batch3dmot_b200 import synth
    gs = []
    for i in range(scenes):
        s = SEED + 1000 * rank + i
        gs.append(synth.add_labels(synth.add_modalities(synth.scene_graph(seed=s), s, raw=False), s))
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


def cpu_reference_run(steps, warmup, n_scenes):
    """The reference's CPU op stream (oracle port, faithful mode) on the host cores."""
    from oracle import ref_restated as R
    from batch3dmot_b200.clr_att_gnn import GNN
    torch.set_num_threads(os.cpu_count())
    data = make_batch(0, n_scenes)
    torch.manual_seed(SEED)
    sd = GNN(None, None, None).state_dict()
    params = {k: v.clone().requires_grad_(not k.startswith("knn_conv")) for k, v in sd.items()}
    mha = R.build_mha(params)
    E = data.edge_index.size(1)
    for _ in range(warmup):
        R.cpu_train_step(params, data, mha)
    t0 = time.perf_counter()
    for _ in range(steps):
        R.cpu_train_step(params, data, mha)
    dt = (time.perf_counter() - t0) / steps
    return E / dt, dt * 1e3, E, torch.get_num_threads()


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
        steps, warm = max(1, min(a.steps, 5)), max(1, min(a.warmup, 2))
        v, ms, E, thr = cpu_reference_run(steps, warm, a.cpu_scenes)
        print(json.dumps({
            "impl": "reference", "metric": "gnn_edges_per_s_fwd_bwd", "value": v, "unit": "edges/s", "n_gpus": a.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": v, "unit": "edges/s", "cores": thr, "kind": "port",
                             "sample": f"{a.cpu_scenes} scene graph(s), {E} edges per step, {steps} steps; reference op "
                                       "stream restated in torch (PyG/torch_scatter are not installable)"},
            "e2e": {"value": v, "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch.distributed as dist
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from batch3dmot_b200 import _lib, ops, build
    from batch3dmot_b200.clr_att_gnn import GNN
    from batch3dmot_b200.parallel import Trainer
    if local == 0:
        build.build()          # no-op when the in-tree libb3d.so is up to date (one rank per node builds)
    if world > 1:
        dist.barrier()
    _lib.lib()   # fail loudly if the CUDA library is missing

    host = make_batch(rank, a.scenes)
    E, N = host.edge_index.size(1), host.num_nodes
    keys = ["pose_feats", "edge_index", "edge_attr", "x_img", "pointnet_out", "radarnet_out", "m_lidar", "m_radar",
            "y", "edge_weights", "node_timestamps"]
    pinned = {k: getattr(host, k).pin_memory() for k in keys}
    h2d_bytes = sum(t.numel() * t.element_size() for t in pinned.values())

    def to_device():
        from types import SimpleNamespace
        return SimpleNamespace(**{k: t.to(dev, non_blocking=True) for k, t in pinned.items()}, num_nodes=N)

    e_tot = torch.tensor([E], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(e_tot)
    E_global = int(e_tot.item())

    torch.manual_seed(SEED)
    model = GNN(None, None, None).to(dev)
    ops.set_precision(a.precision)
    trainer = Trainer(model, batch_size=2)

    def fwd_kwargs(d):
        return dict(x_img=d.x_img, pointnet_out=d.pointnet_out, radarnet_out=d.radarnet_out, lidar_mask=d.m_lidar,
                    radar_mask=d.m_radar)

    # ---- device-resident timing ("value")
    d = to_device()
    d._b3d_graph = ops.Graph(d.edge_index, N)
    kw = fwd_kwargs(d)
    for _ in range(a.warmup):
        trainer.step(d, global_edges=E_global, **kw)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    barrier()
    _lib.reset_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        ev0.record()
        for _ in range(a.steps):
            loss = trainer.step(d, global_edges=E_global, **kw)
        ev1.record()
        barrier()
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
        for _ in range(2):
            model(d, **kw)
        barrier()
        ev0.record()
        for _ in range(a.steps):
            model(d, **kw)
        ev1.record()
        barrier()
    model.train()
    msf = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    if world > 1:
        dist.all_reduce(msf, op=dist.ReduceOp.MAX)
    fwd_value = E_global / (float(msf.item()) / a.steps * 1e-3)

    # ---- end-to-end through the public API with host buffers ("e2e"): every step copies ITS inputs
    # from pinned host memory (on a copy stream, overlapped with the previous step's kernels), builds
    # the CSR from the fresh edge_index, runs the step and reads the loss back to the host.
    # Two preallocated device staging sets (double buffer): no allocation inside the loop; a set is
    # overwritten only after the step that read it has finished (event), and the in-place copy bumps the
    # tensor version, so the CSR cache misses and the tables are rebuilt from the fresh edge_index.
    copy_stream = torch.cuda.Stream(device=dev)
    from types import SimpleNamespace
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
            l = trainer.step(dd, global_edges=E_global, **fwd_kwargs(dd))   # builds CSR from edge_index
            done[i & 1] = torch.cuda.Event()
            done[i & 1].record()
            last = float(l.item())                          # D2H of the loss (host sync every step)
        return last

    e2e_loop(max(3, a.warmup))     # warm-up: both staging buffer sets and the per-step CSR tables get allocated
    barrier()
    ev0.record()
    e2e_loop(a.steps)
    ev1.record()
    barrier()
    ms2 = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_value = E_global / (float(ms2.item()) / a.steps * 1e-3)

    # ---- roofline of the dominant kernel
    hbm, tf_burst, tf_sust, src = peaks()
    def time_kernel(fn, reps=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(reps):
            fn()
        ev1.record()
        torch.cuda.synchronize()
        return ev0.elapsed_time(ev1) / reps * 1e-3

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
                "traffic": 848_617_984,   # dram read+write bytes per launch (502,058,240 + 346,559,744), ncu --set full, same shape
                "traffic_source": "profiles/r1_roofline_kernel.md (algorithmic bytes: E*(512+384)*2 = 877,743,104)",
                "peak_source": src + " HBM copy bandwidth (burst); kernel timed alone with CUDA events",
                "rows": Er, "us_per_launch": t * 1e6,
                "tensor_view": {"achieved_tflops": ach_t, "peak_tflops": tf_burst, "frac": ach_t / tf_burst}}
    else:
        # fp32 exact path: the dominant launch is k_linear on edge_update layer 0 ([E,320] -> 256, gathered)
        mp = model.message_passing
        lin = mp.edge_update[0]
        x = torch.randn(N, 96, device=dev); e = torch.randn(E, 64, device=dev); att = torch.randn(E, 64, device=dev)
        g = d._b3d_graph
        items = [(x, g.dst32, None, 0), (x, g.src32, None, 0), (e, None, None, 0), (att, None, None, 0)]
        out = torch.empty(E, 256, device=dev)
        t = time_kernel(lambda: ops.linear_raw(items, lin.weight, lin.bias, E, 1, out=out))
        ach = 2.0 * E * 320 * 256 / t / 1e12
        roof = {"kernel": "k_linear (fp32 FFMA; edge_update layer 0, gathered [E,320]x[320,256])", "bound": "tensor",
                "achieved": ach, "peak": tf_burst, "unit": "TFLOP/s", "frac": ach / tf_burst, "traffic": None,
                "peak_source": src + " bf16 dense burst; kernel timed alone",
                "note": "fp32-exact SIMT path: runs on the FFMA pipe, not the tensor pipe"}

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
            ent = agg.setdefault(cls, [0, 0.0, 0])
            ent[0] += n; ent[1] += ms_k; ent[2] += nb
        roof_step = {cls: {"launches": n, "ms": ms_k, "algorithmic_gb": nb / 1e9, "achieved_gbs": nb / ms_k / 1e6,
                           "frac_of_hbm_peak": nb / ms_k / 1e6 / hbm}
                     for cls, (n, ms_k, nb) in agg.items() if ms_k > 0}

    line = {"metric": "gnn_edges_per_s_fwd_bwd", "value": value, "unit": "edges/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if a.precision == "fp32" else "bf16", "data": "synthetic",
            "config": dict(config, edges_per_step=E_global, nodes_per_gpu=N),
            "e2e": {"value": e2e_value, "unit": "edges/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4},
            "gpu_launches": launches, "clocks": clk.summary(), "roofline": roof, "roofline_step": roof_step,
            "algorithmic_tflops": value * MM_FWDBWD_FLOPS_PER_EDGE / 1e12, "loss": float(loss.item()),
            "forward_only": {"value": fwd_value, "unit": "edges/s", "ms_per_step": E_global / fwd_value * 1e3,
                             "note": "inference forward of the same batch under torch.no_grad()"},
            "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 2**30}
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        v, cms, cE, thr = cpu_reference_run(3, 1, a.cpu_scenes)
        line["cpu_baseline"] = {"value": v, "unit": "edges/s", "cores": thr, "kind": "port",
                                "sample": f"{a.cpu_scenes} scene graph(s), {cE} edges per step, 3 steps "
                                          f"({cms:.0f} ms/step); reference op stream restated in torch"}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
