"""CPU (gloo, world_size 2) tests of the scene-sharded data-parallel host logic: LPT partition,
flat gradient buffer + SUM all-reduce, and the loss re-weighting that makes an R-rank step equal a
1-rank step on the union batch. The model arithmetic here is the CPU oracle — the
CUDA kernels are covered by the -m gpu tests; this file checks only the distributed plumbing."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from batch3dmot_b200 import synth
from batch3dmot_b200.parallel import FlatParams, allreduce_sum_, lpt_partition


def test_lpt_partition_balances_and_is_deterministic():
    costs = [61000, 5000, 30000, 30000, 8000, 52000, 1000, 44000, 44000, 12000]
    bins = lpt_partition(costs, 4)
    assert sorted(i for b in bins for i in b) == list(range(len(costs)))
    loads = [sum(costs[i] for i in b) for b in bins]
    assert max(loads) - min(loads) <= max(costs)            # LPT guarantee-ish on this instance
    assert max(loads) <= 1.34 * sum(costs) / 4              # within 4/3 of the ideal makespan
    assert bins == lpt_partition(costs, 4)
    assert lpt_partition([], 3) == [[], [], []]
    assert lpt_partition([5], 2) == [[0], []]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _tiny_model():
    torch.manual_seed(5621)
    return torch.nn.Sequential(torch.nn.Linear(4, 16), torch.nn.ReLU(), torch.nn.Linear(16, 1))


def _loss(model, g):
    out = torch.sigmoid(model(g.edge_attr.float()))
    w = g.edge_weights
    return torch.nn.functional.binary_cross_entropy(out.view(-1), g.y.float(), weight=w) / 2


def _worker(rank, world, port, scenes, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    graphs = [synth.add_labels(synth.scene_graph(seed=s, T=5, nodes_per_frame=8, k=6), s, p_pos=0.2) for s in scenes]
    costs = [g.edge_index.size(1) for g in graphs]
    mine = lpt_partition(costs, world)[rank]
    local = synth.collate([graphs[i] for i in mine])
    model = _tiny_model()
    fp = FlatParams(model)
    e_local = torch.tensor([local.edge_index.size(1)])
    e_glob = e_local.clone()
    dist.all_reduce(e_glob)
    fp.zero_grad()
    loss = _loss(model, local) * (e_local.item() / e_glob.item())     # Trainer.step's re-weighting
    loss.backward()
    assert all(p.grad.data_ptr() == fp.grad[o:o + p.numel()].data_ptr()
               for p, o in zip(fp.params, [0, 64, 80, 96]))           # grads landed in the flat buffer
    allreduce_sum_(fp.grad)
    q.put((rank, fp.grad.clone(), int(e_glob.item())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_step_equals_single_rank_union_batch():
    scenes = [11, 12, 13, 14, 15]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, scenes, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-rank reference on the union batch
    graphs = [synth.add_labels(synth.scene_graph(seed=s, T=5, nodes_per_frame=8, k=6), s, p_pos=0.2) for s in scenes]
    union = synth.collate(graphs)
    model = _tiny_model()
    fp = FlatParams(model)
    fp.zero_grad()
    _loss(model, union).backward()
    for rank, g, e_glob in results:
        assert e_glob == union.edge_index.size(1)
        assert torch.allclose(g, fp.grad, rtol=1e-5, atol=1e-8), f"rank {rank} gradient differs from the union-batch step"
    assert torch.equal(results[0][1], results[1][1])                  # identical on every rank after the all-reduce


def test_flat_params_views_track_updates():
    model = _tiny_model()
    ref = [p.detach().clone() for p in model.parameters()]
    fp = FlatParams(model)
    assert fp.numel == sum(p.numel() for p in ref)
    for p, r in zip(model.parameters(), ref):
        assert torch.equal(p, r)
    fp.flat.add_(1.0)                                                 # an optimiser writing the flat buffer
    for p, r in zip(model.parameters(), ref):
        assert torch.equal(p, r + 1.0)
