"""Scene-sharded data parallelism. Scenes / window graphs are independent
disjoint graphs (predict.py:172-190 builds one Data per window; the graph construction only links
same-category nodes inside a window), so:
  * inference shards graphs across ranks with NO data-path collective;
  * training has exactly one exchange step: a SUM all-reduce of one flat fp32 gradient buffer
    (1,319,697 floats = 5.28 MB for the multimodal model), then the same Adam step on every rank.
The reference trains the GNN on a single device (train.py:45); its only DDP precedent is the
encoder trainer (training/train_resnet_ae_ddp.py:125-175)."""
import torch
import torch.distributed as dist


def lpt_partition(costs, world_size):
    """Longest-processing-time bin packing: returns `world_size` lists of item ids, balancing
    the summed cost (edge counts vary ~10x between scenes). Deterministic: ties by id."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    loads = [0] * world_size
    bins = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda j: (loads[j], j))
        bins[r].append(i)
        loads[r] += costs[i]
    for b in bins:
        b.sort()
    return bins


class FlatParams:
    """Re-homes a module's trainable parameters (and their .grad) as views of two flat fp32
    buffers, so weight-gradient kernels accumulate straight into ONE contiguous all-reduce
    payload and the optimiser is one kernel."""

    def __init__(self, module, params=None):
        self.params = [p for p in module.parameters() if p.requires_grad] if params is None else list(params)
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.empty(n, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            k = p.numel()
            self.flat[off:off + k].copy_(p.detach().reshape(-1))
            p.data = self.flat[off:off + k].view_as(p)
            p.grad = self.grad[off:off + k].view_as(p)
            off += k
        self.numel = n

    def zero_grad(self):
        self.grad.zero_()
        off = 0
        for p in self.params:      # re-attach in case autograd replaced .grad
            k = p.numel()
            if p.grad is None or p.grad.data_ptr() != self.grad[off:off + k].data_ptr():
                p.grad = self.grad[off:off + k].view_as(p)
            off += k


def allreduce_sum_(flat_grad):
    """The single training collective: NCCL SUM all-reduce over NVLink (gloo in CPU tests)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
    return flat_grad


class Trainer:
    """fwd + weighted BCE + bwd + gradient all-reduce + Adam, mirroring train.py:126-160
    (Adam lr 1e-4, weight_decay 1e-4, betas .9/.999; loss / batch_size).

    Parameters that never receive a gradient (knn_conv.* while its result is discarded, pose_gnn.py:80)
    are left untouched, exactly like torch.optim.Adam skips parameters whose .grad is None: the first
    step runs with .grad = None everywhere, and only the parameters autograd reached are re-homed into
    the flat buffers that the all-reduce and the Adam kernel operate on."""

    def __init__(self, model, lr=1e-4, weight_decay=1e-4, betas=(0.9, 0.999), eps=1e-8, batch_size=2,
                 from_logits=False, loss="bce", focal_alpha=0.25, focal_gamma=2.0, data_parallel=True):
        from . import ops
        self.ops = ops
        self.model = model
        self.fp = self.m = self.v = None
        self.lr, self.wd, self.betas, self.eps = lr, weight_decay, betas, eps
        self.batch_size, self.from_logits = batch_size, from_logits
        assert loss in ("bce", "focal")      # "focal": BASELINE config 5's alternative edge loss (not in the reference)
        self.loss, self.focal = loss, (focal_alpha, focal_gamma)
        self.step_no = 0
        self._step_dev = None
        # data_parallel=False: a rank-local trainer inside a multi-rank job (no gradient all-reduce, no loss re-weighting)
        self.data_parallel = data_parallel

    def _loss(self, data, global_edges, fwd_kwargs):
        ops = self.ops
        out, _ = self.model(data, **fwd_kwargs)
        if self.loss == "focal":
            loss = ops.focal_loss(out, data.y, getattr(data, "edge_weights", None), batch_size=self.batch_size,
                                  alpha=self.focal[0], gamma=self.focal[1], from_logits=self.from_logits)
        else:
            loss = ops.bce_loss(out, data.y, getattr(data, "edge_weights", None), batch_size=self.batch_size,
                                from_logits=self.from_logits)
        world = dist.get_world_size() if (self.data_parallel and dist.is_available() and dist.is_initialized()) else 1
        if global_edges is not None:
            loss = loss * (out.size(0) / float(global_edges))
        elif world > 1:
            loss = loss / world          # SUM all-reduce of per-rank mean losses -> mean over ranks (DDP convention)
        return loss

    def _first_step(self, data, global_edges, fwd_kwargs):
        params = [p for p in self.model.parameters() if p.requires_grad]
        for p in params:
            p.grad = None
        loss = self._loss(data, global_edges, fwd_kwargs)
        loss.backward()
        used = [p for p in params if p.grad is not None]
        first = [p.grad for p in used]
        self.fp = FlatParams(self.model, params=used)
        for p, g in zip(used, first):
            p.grad.copy_(g)
        self.m = torch.zeros_like(self.fp.flat)
        self.v = torch.zeros_like(self.fp.flat)
        return loss

    def step(self, data, global_edges=None, **fwd_kwargs):
        """One optimisation step on this rank's shard. With `global_edges` (sum of E over ranks)
        the local mean loss is re-weighted so that the SUM all-reduce yields exactly the gradient
        of the mean loss over the union batch; without it the per-rank mean losses are averaged."""
        ops = self.ops
        if self.fp is None:
            loss = self._first_step(data, global_edges, fwd_kwargs)
        else:
            self.fp.zero_grad()
            loss = self._loss(data, global_edges, fwd_kwargs)
            with ops.grads_into_params():      # weight-gradient kernels accumulate straight into the flat buffer
                loss.backward()
        if self.data_parallel:
            allreduce_sum_(self.fp.grad)
        self.step_no += 1
        if self._step_dev is not None:      # graph-capturable form: the step number lives on the device
            ops.adam_step_dev(self.fp.flat, self.fp.grad, self.m, self.v, self.lr, self.betas, self.eps, self.wd,
                              self._step_dev)
        else:
            ops.adam_step(self.fp.flat, self.fp.grad, self.m, self.v, self.lr, self.betas, self.eps, self.wd,
                          self.step_no)
        ops.invalidate_weight_cache()      # raw-pointer update: packed bf16 copies are stale
        return loss.detach()

    def capture(self, data, global_edges=None, warmup=2, **fwd_kwargs):
        """Capture one whole training step on `data` (static shapes and addresses) in a CUDA graph and return
        `replay() -> loss tensor`. For the launch-bound small-batch regime (the reference trains with 2 window
        graphs per step, cl_config.yaml:99: a few hundred short kernels per step): the graph removes the Python /
        ctypes / autograd dispatch between them. Everything the step touches is static: parameters, gradients and
        Adam moments are the flat buffers, the weight packs are re-made inside the graph from the updated
        parameters, the step number is a device counter, and NCCL all-reduces are capturable."""
        dev = self.model.parameters().__next__().device
        if self.fp is None:
            self.step(data, global_edges, **fwd_kwargs)             # eager: decides which parameters train
        if self._step_dev is None:
            self._step_dev = torch.full((1,), self.step_no, dtype=torch.int32, device=dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.step(data, global_edges, **fwd_kwargs)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            loss = self.step(data, global_edges, **fwd_kwargs)

        def replay():
            graph.replay()
            self.step_no += 1
            return loss
        replay.graph = graph
        return replay
