"""Host-side operators over the libb3d C ABI: thin wrappers plus torch.autograd.Functions so
the drop-in layers train with the reference's own `loss.backward(); optimizer.step()`
(train.py:157-160). Every FLOP on this path runs in a libb3d kernel; torch provides device
memory, streams and autograd bookkeeping only."""

import os
import weakref

import torch

from . import _lib as L


# ----------------------------------------------------------------------------- per-op timing (debug)
_OPTIME = None      # dict signature -> [count, ms, bytes] when enabled by optime_begin()


def optime_begin():
    global _OPTIME
    _OPTIME = {}


def optime_end():
    global _OPTIME
    r, _OPTIME = _OPTIME, None
    return r


class _timed:
    """Debug helper: CUDA-event time of one wrapped launch, keyed by an op signature (synchronises!)."""

    def __init__(self, sig, nbytes):
        self.sig, self.nbytes = sig, nbytes

    def __enter__(self):
        if _OPTIME is not None:
            self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *a):
        if _OPTIME is not None:
            self.e1.record()
            torch.cuda.synchronize()
            ent = _OPTIME.setdefault(self.sig, [0, 0.0, 0])
            ent[0] += 1; ent[1] += self.e0.elapsed_time(self.e1); ent[2] += self.nbytes


# ----------------------------------------------------------------------------- graph tables
class NodeIndex:
    """One endpoint of the edge list: int32 ids per edge plus the CSR that groups edges by
    that endpoint (rowptr [n+1], stable perm [E])."""
    __slots__ = ("idx", "rowptr", "perm", "n", "sorted")

    def __init__(self, idx, rowptr, perm, n, is_sorted=False):
        self.idx, self.rowptr, self.perm, self.n, self.sorted = idx, rowptr, perm, n, is_sorted


class Graph:
    """CSC (by target) + CSR (by source) of an edge_index, built once on device by the radix
    sort in csr.cu (replaces the per-call atomics of torch_scatter, pose_gnn.py:190-191)."""

    def __init__(self, edge_index, num_nodes, check=False):
        assert edge_index.is_cuda and edge_index.dtype == torch.int64 and edge_index.dim() == 2 \
            and edge_index.size(0) == 2, "edge_index must be a CUDA int64 [2,E] tensor"  # __check_input__
        ei = edge_index.contiguous()
        dev = ei.device
        E, N = ei.size(1), int(num_nodes)
        i32 = dict(dtype=torch.int32, device=dev)
        self.E, self.N = E, N
        self.src32, self.dst32 = torch.empty(E, **i32), torch.empty(E, **i32)
        self.rowptr_dst, self.perm_dst = torch.empty(N + 1, **i32), torch.empty(E, **i32)
        self.rowptr_src, self.perm_src = torch.empty(N + 1, **i32), torch.empty(E, **i32)
        self.status = torch.zeros(1, **i32)
        lib = L.lib()
        wsb = lib.b3d_csr_workspace_bytes(E, N)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        L.check(lib.b3d_csr_build(L.ptr(ei), E, N, L.ptr(self.src32), L.ptr(self.dst32),
                                  L.ptr(self.rowptr_dst), L.ptr(self.perm_dst), L.ptr(self.rowptr_src),
                                  L.ptr(self.perm_src), L.ptr(ws), wsb, L.ptr(self.status), L.stream()),
                "b3d_csr_build")
        if check and int(self.status.item()) != 0:   # optional (syncs)
            raise IndexError("edge_index holds node ids outside [0, num_nodes)")
        self.by_dst = NodeIndex(self.dst32, self.rowptr_dst, self.perm_dst, N)
        self.by_src = NodeIndex(self.src32, self.rowptr_src, self.perm_src, N)
        self._deg = {}

    def degree_block(self, nidx, dtype=torch.float32):
        """[N,8]: the number of edges whose endpoint is the node, split as column 0 = deg % 256 and
        column 1 = deg // 256 (both exact in bf16; the matching weight columns are b and 256 b), columns
        2-7 zero: an 8-wide block so it can be a segment of a tensor-core layer input."""
        key = (id(nidx), dtype)
        if key not in self._deg:
            deg = nidx.rowptr[1:] - nidx.rowptr[:-1]
            d = torch.zeros((self.N, 8), dtype=torch.float32, device=nidx.rowptr.device)
            d[:, 0] = (deg % 256).to(torch.float32)
            d[:, 1] = (deg // 256).to(torch.float32)
            self._deg[key] = d.to(dtype)
        return self._deg[key]


_graph_cache = {}


def graph_of(edge_index, num_nodes):
    """Cache keyed by the edge_index storage, so the sort really is one-time per graph."""
    key = (edge_index.data_ptr(), tuple(edge_index.shape), edge_index._version, int(num_nodes))
    g = _graph_cache.get(key)
    if g is None:
        # small FIFO: a stream of fresh edge_index tensors (one per step in an end-to-end loop) must not pin
        # dozens of E-sized tables, or every step's tables come from a fresh cudaMalloc instead of the
        # allocator's free list
        while len(_graph_cache) >= 4:
            _graph_cache.pop(next(iter(_graph_cache)))
        g = _graph_cache[key] = Graph(edge_index, num_nodes)
        g._keepalive = edge_index
    return g


# ----------------------------------------------------------------------------- precision mode
# "fp32" : the 1e-4 parity mode. Dense layers run on the tensor cores in SPLIT-bf16 arithmetic (every fp32 operand
#          as hi + lo bf16 parts, three tcgen05.mma per K step, fp32 accumulation: ~2^-16 relative per product);
#          shapes the tiles do not take (tiny M, ragged widths, operand masks) use the FFMA kernel.
# "exact": FFMA kernels everywhere (plain fp32 arithmetic; kernel-level tests against float64).
# "bf16" : bf16 operands and bf16 storage between layers, fused chains (2e-2 mode, the benchmarked path).
_PRECISION = "fp32"
_TC_MIN_ROWS = 256
_packed = {}


def set_precision(p):
    """Select the arithmetic of the dense layers: 'fp32' (exact mode) or 'bf16' (tensor-core tiles
    with fp32 accumulation; layers whose shapes do not fit the tile constraints stay fp32)."""
    global _PRECISION
    assert p in ("fp32", "exact", "bf16")
    _PRECISION = p


def get_precision():
    return _PRECISION


def preprojected():
    """The node-side first-layer blocks are applied per NODE in the bf16 mode and in the tensor-core
    1e-4 mode; the "exact" (FFMA) mode keeps the reference's per-edge concatenation."""
    return _PRECISION == "bf16" or (_PRECISION == "fp32" and FEATURES["split_tc"])


def store_dtype():
    """Storage type of edge / node tensors between kernels in the pre-projected formulation."""
    return torch.bfloat16 if _PRECISION == "bf16" else torch.float32


_weight_epoch = 0        # bumped by invalidate_weight_cache(); part of every derived-weight key


def invalidate_weight_cache():
    """Packed bf16 weights are cached per live tensor object and version; call after updating weights through
    raw pointers (b3d_adam_step does not bump torch's version counter)."""
    global _weight_epoch
    _weight_epoch += 1
    _packed.clear()
    _derived.clear()
    _chain_packs.clear()


# Weight gradients straight into .grad. Inside `with grads_into_params():` (parallel.Trainer wraps its loss.backward()
# in it once the flat gradient buffer exists) the weight-gradient kernels of _FusedMLP ACCUMULATE into the .grad of a leaf
# parameter — or into the matching row / column window of it when the layer's weight is a slice view of a parameter —
# and autograd receives no gradient for that input: no fresh dW tensor, no AccumulateGrad addition per use of a shared
# weight (6 iterations x ~16 tensors), no zero-filled full-size gradient per sliced block. Same additions in the same
# order as autograd's accumulation, so the result is unchanged. Off by default: torch.autograd.grad(), hooks and
# .grad = None flows keep autograd's own semantics.
_GRAD_SINK = False


class grads_into_params:
    def __enter__(self):
        global _GRAD_SINK
        self.prev, _GRAD_SINK = _GRAD_SINK, True

    def __exit__(self, *a):
        global _GRAD_SINK
        _GRAD_SINK = self.prev


def _grad_target(w):
    """The fp32 window of a leaf parameter's .grad that corresponds to `w` (the parameter itself or a view of it)."""
    if not _GRAD_SINK or w is None or w.dtype != torch.float32:
        return None
    base = w._base if w._base is not None else w
    if not (base.is_leaf and base.requires_grad):      # (.grad of a non-leaf tensor must not even be looked at)
        return None
    g = base.grad
    if g is None or g.dtype != torch.float32 or g.device != w.device or g.shape != base.shape or g.stride() != base.stride():
        return None
    if w is base:
        return g
    return torch.as_strided(g, w.size(), w.stride(), g.storage_offset() + w.storage_offset() - base.storage_offset())


_USE_TMA = True          # dense bf16 operands go through the TMA-fed persistent kernel
_USE_BITS = True         # bf16 chains keep ReLU masks as sign bits (B3D_BITS) for the backward pass


def _packed_weight_segs(W, transpose, widths):
    """Row-major TMA pack with every input segment padded to whole 64-column chunks (gathered / ragged segments)."""
    key = (id(W), tuple(W.shape), W.stride(0), bool(transpose), "segs", tuple(widths))
    ent = _packed.get(key)
    if ent is not None:
        ref, ver, ptr, wp = ent
        if ref() is W and ver == W._version and ptr == W.data_ptr():
            return wp
    if len(_packed) > 512:
        _packed.clear()
    n_log = W.size(1) if transpose else W.size(0)
    lib = L.lib()
    wa = L.int_array(list(widths))
    wp = torch.empty(int(lib.b3d_tma_packed_bytes_segs(n_log, wa, len(widths))), dtype=torch.uint8, device=W.device)
    L.check(lib.b3d_tma_pack_weights_segs(L.ptr(W), W.stride(0), n_log, wa, len(widths), int(transpose), L.ptr(wp),
                                          L.stream()), "b3d_tma_pack_weights_segs")
    _packed[key] = (weakref.ref(W), W._version, W.data_ptr(), wp)
    return wp


def _packed_weight(W, transpose, rowmajor=False, split=False):
    """bf16 pack of W for the tensor-core kernels. An entry belongs to ONE live tensor object: it holds a weak
    reference and is used only while that very object is alive at the same version. (A key made of
    (data_ptr, _version) alone would return the old pack for a temporary that the caching allocator placed
    at a recycled address, or for the parameters of a new model allocated where a freed one lived.)"""
    key = (id(W), tuple(W.shape), W.stride(0), bool(transpose), rowmajor, split)
    ent = _packed.get(key)
    if ent is not None:
        ref, ver, ptr, wp = ent
        if ref() is W and ver == W._version and ptr == W.data_ptr():
            return wp
    if len(_packed) > 512:
        _packed.clear()
    n_log, k_log = (W.size(1), W.size(0)) if transpose else (W.size(0), W.size(1))
    lib = L.lib()
    nbytes, pack = (lib.b3d_tma_packed_bytes, lib.b3d_tma_pack_weights) if rowmajor else \
        (lib.b3d_tc_packed_bytes, lib.b3d_tc_pack_weights)
    assert not (split and rowmajor)
    wp = torch.empty(nbytes(n_log, k_log) * (4 if split else 1), dtype=torch.uint8, device=W.device)
    L.check(pack(L.ptr(W), W.stride(0), n_log, k_log, int(transpose) | (2 if split else 0), L.ptr(wp), L.stream()),
            "pack_weights")
    _packed[key] = (weakref.ref(W), W._version, W.data_ptr(), wp)
    return wp


_derived = {}


def derived_weights(tag, sources, build):
    """Tensors DERIVED from parameters (stacked / sliced weight blocks of the pre-projected formulation),
    rebuilt only when a source parameter changes: the key holds every source's identity, address and version
    plus the invalidation epoch, and the entry keeps weak references to the sources, so neither an optimiser
    step, load_state_dict, nor a new model at a recycled address can hit a stale entry. Under autograd the
    derived tensors must stay differentiable functions of the parameters, so nothing is cached there."""
    if torch.is_grad_enabled() and any(s.requires_grad for s in sources):
        return build()
    key = (tag, _weight_epoch) + tuple((id(s), s.data_ptr(), s._version) for s in sources)
    ent = _derived.get(key)
    if ent is not None and all(r() is s for r, s in zip(ent[0], sources)):
        return ent[1]
    if len(_derived) > 64:
        _derived.clear()
    val = build()
    _derived[key] = ([weakref.ref(s) for s in sources], val)
    return val


_chain_packs = {}
# Switches for the kernel families added in round 2 (B3D_FEATURES=all|none|comma list overrides the defaults):
#   split_tc   tf32 x3 tensor-core tiles as the arithmetic of the "fp32" (1e-4) mode (else: FFMA kernels). ON:
#              every 1e-4 test passes with it (tests/test_gpu_models.py, test_gpu_split.py) and it is 2.8x faster.
#   window_knn graph-construction k-NN kernel for CUDA tensors (window_knn.cu). ON.
#   gather_tma bf16 mode: the node-feature operands of the edge layers are row-gathered by the TMA unit
#              (tile::gather4) straight into the tensor-core tile — the reference's un-projected cat[x_i, x_j, e, att]
#              form — instead of arriving as epilogue addends of per-node pre-projections.
#   edge_block the edge side of a message-passing iteration as ONE autograd node with an explicit backward
#              (ops._MPEdgeBlockG): same kernels in the forward pass, a K-concatenated de' GEMM instead of two
#              GEMMs + a 3-way sum kernel in the backward pass. Same speed as the default (86.2 vs 86.1 ms/step), 10 GB
#              less memory at 64 scenes: OFF (nothing to gain in time).
#   narrow_split bf16 mode: the widest layer of the two narrow chains (classifier 64 -> 32, edge encoder 32 -> 64) as a
#              tensor-core layer, the rest in the narrow kernel (_narrow_split). ON.
#   bf16_inputs bf16 mode: dense fp32 inputs of a tensor-core chain are rounded to bf16 once (not per tile). ON.
#   chain      fused MLP chains / edge blocks (chain_tc.cu). Validated (tests/test_gpu_chain.py runs it whatever this
#              switch says) but OFF in the model path: with one 128-row tile in flight per SM the fused kernel is
#              bound by the same epilogue work as the per-layer kernels plus the layer-to-layer hand-over latency,
#              and measured no faster (profiles/r2_chain_kernel.md), so the per-layer TMA kernels stay the default.
_FEATURE_DEFAULTS = {"chain": False, "split_tc": True, "window_knn": True, "gather_tma": False, "edge_block": False,
                     "narrow_split": True, "bf16_inputs": True}


def _read_features():
    env = os.environ.get("B3D_FEATURES")
    if env is None:
        return dict(_FEATURE_DEFAULTS)
    on = set(_FEATURE_DEFAULTS) if env.strip() == "all" else {t.strip() for t in env.split(",") if t.strip() not in ("", "none")}
    assert on <= set(_FEATURE_DEFAULTS), f"unknown feature in B3D_FEATURES={env!r}"
    return {k: k in on for k in _FEATURE_DEFAULTS}


FEATURES = _read_features()
_USE_CHAIN = FEATURES["chain"]   # multi-layer bf16 chains run as ONE fused launch (chain_tc.cu)


def _wkey(w):
    """Identity of a weight (or of a column/row slice VIEW of one) for pack caches: the base tensor object,
    its address and version, plus the view geometry. Slices are re-created on every forward, their base is not."""
    base = w._base if w._base is not None else w
    return (id(base), base.data_ptr(), base._version, w.storage_offset(), tuple(w.shape), w.stride()), base


def chain_run(inputs, specs, idx0, idx1, M):
    """One fused launch of a layer program (b3d_chain_run). inputs: 1-2 dense bf16 [M, w] tensors.
    specs: per layer dict(W fp32 [N,K] or its transpose with transpose=True, src, act, bias, adds=[(bf16 tensor, sel)],
    out, bits_out, bits_in). Returns False (nothing launched) when the chain does not fit the kernel's plan."""
    lib = L.lib()
    nl = len(specs)
    k_in = sum(t.size(1) for t in inputs)
    arr = (L.ChainLayer * nl)()
    for l, sp in enumerate(specs):
        W = sp["W"]
        n, k = (W.size(1), W.size(0)) if sp.get("transpose") else (W.size(0), W.size(1))
        c = arr[l]
        c.K, c.N, c.src, c.act = k, n, sp.get("src", l - 1), sp.get("act", L.ACT_NONE)
        adds = sp.get("adds") or []
        c.nadd = len(adds)
        for t, (at, sel) in enumerate(adds):
            assert at.dtype == torch.bfloat16 and at.size(1) == n and at.stride(1) == 1
            c.add_ptr[t], c.add_ld[t], c.add_idx[t] = at.data_ptr(), at.stride(0), sel
        b = sp.get("bias")
        c.bias = b.data_ptr() if b is not None else None
        o = sp.get("out")
        if o is not None:
            assert o.dtype == torch.bfloat16 and o.shape == (M, n) and o.stride(1) == 1
            c.out, c.ldo = o.data_ptr(), o.stride(0)
        for name in ("bits_out", "bits_in"):
            bt = sp.get(name)
            if bt is not None:
                assert bt.shape == ((n + 31) // 32, M) and bt.is_contiguous()
                setattr(c, name, bt.data_ptr())
    if not lib.b3d_chain_supported(arr, nl, k_in):
        return False
    for sp in specs:      # the epilogue moves 32 bytes per lane: rows of addends / outputs must be 32-byte aligned
        for t in [at for at, _ in (sp.get("adds") or [])] + ([sp["out"]] if sp.get("out") is not None else []):
            if t.data_ptr() % 32 or t.stride(0) % 16:
                return False
    keys, bases = zip(*[_wkey(sp["W"]) for sp in specs])
    key = (_weight_epoch, k_in, tuple(keys), tuple(bool(sp.get("transpose")) for sp in specs))
    ent = _chain_packs.get(key)
    if ent is None or not all(r() is b for r, b in zip(ent[0], bases)):
        if len(_chain_packs) > 64:
            _chain_packs.clear()
        buf = torch.empty(int(lib.b3d_chain_packed_bytes(arr, nl, k_in)), dtype=torch.uint8, device=inputs[0].device)
        for l, sp in enumerate(specs):
            W = sp["W"]
            assert W.dtype == torch.float32 and W.stride(1) == 1
            L.check(lib.b3d_chain_pack_weights(arr, nl, k_in, l, L.ptr(W), W.stride(0), int(bool(sp.get("transpose"))),
                                               L.ptr(buf), L.stream()), "b3d_chain_pack_weights")
        ent = _chain_packs[key] = ([weakref.ref(b) for b in bases], buf)
    segs = L.make_segs([(t, None, None, 0) for t in inputs])
    L.check(lib.b3d_chain_run(segs, len(inputs), arr, nl, L.ptr(ent[1]), L.ptr(idx0), L.ptr(idx1), M, L.stream()),
            "b3d_chain_run")
    return True


def _chain_dims_ok(widths):
    return all(w % 64 == 0 and 64 <= w <= 512 for w in widths)


def _tma_ok(items, K, accumulate):
    """Dense bf16 operands (1-2 segments, only the last ragged): the TMA-fed kernel with the plain weight pack.
    With the gather_tma feature also up to 6 segments, row-gathered ones included (two index arrays at most)."""
    if not _USE_TMA or accumulate or K > 1024:
        return False
    general = FEATURES["gather_tma"]
    if len(items) > (6 if general else 2):
        return False
    idxs = []
    for n, (t, idx, mask, _) in enumerate(items):
        if t.dtype != torch.bfloat16 or mask is not None or not _al16(t) or t.size(1) % 8:
            return False
        if idx is not None:
            if not general:
                return False
            if not any(idx is j for j in idxs):
                idxs.append(idx)
        elif not general and n + 1 < len(items) and t.size(1) % 64:
            return False
    return len(idxs) <= 2 and sum((t.size(1) + 63) // 64 * 64 for t, _, _, _ in items) <= 1024


def _tma_dense_ok(items, K, accumulate):
    """1-2 dense bf16 segments, only the last ragged (weight-gradient TMA kernel, fused chains)."""
    if not _USE_TMA or accumulate or len(items) > 2 or K > 1024:
        return False
    for n, (t, idx, _, _) in enumerate(items):
        if t.dtype != torch.bfloat16 or idx is not None or (n + 1 < len(items) and t.size(1) % 64):
            return False
    return True


def _tma_needs_seg_pack(items):
    return any(idx is not None for _, idx, _, _ in items) or any(t.size(1) % 64 for t, _, _, _ in items[:-1])


def _al16(t):
    q = 8 if t.dtype == torch.bfloat16 else 4
    return t.data_ptr() % 16 == 0 and t.stride(0) % q == 0


def _tc_shapes_ok(items, M, n_out, K):
    """Tile constraints of the tcgen05 kernels (b3d.h): widths % 8, aligned rows, enough work."""
    if M < _TC_MIN_ROWS or n_out < 8 or K < 8:
        return False
    for t, _, mask, _ in items:
        if t.size(1) % 8 or not _al16(t) or mask is not None:
            return False
    return True


# ----------------------------------------------------------------------------- raw wrappers
def _rows(t):
    if t.dim() == 1:
        t = t.unsqueeze(1)
    if t.stride(1) != 1 or t.dtype not in (torch.float32, torch.bfloat16):
        t = t.contiguous().float()
    return t


_DT = {torch.float32: L.F32, torch.bfloat16: L.BF16, torch.float64: L.F64}


def new_relu_bits(M, n_out, device):
    """Sign-bit mask buffer of an [M, n_out] ReLU output: int32 words [ceil(n_out/32), M] (b3d.h B3D_BITS)."""
    return torch.empty(((n_out + 31) // 32, M), dtype=torch.int32, device=device)


def _linear_raw_impl(items, W, bias, M, act=L.ACT_NONE, trans_w=False, out=None, accumulate=False,
               out_mask=None, row_mask=None, n_out=None, tc=None, out_dtype=torch.float32, adds=None,
               bits_out=None, mask_bits=None):
    """bits_out / mask_bits (tensor-core kernels only): sign-bit masks (new_relu_bits) written for this
    layer's output / applied as the producing layer's ReLU mask instead of out_mask."""
    """items: [(tensor, idx32|None, mask|None, mode)]. Returns Y [M, n_out].
    tc=None: use the tensor-core kernel iff the precision mode is bf16 and the shapes fit."""
    segs = L.make_segs(items)
    if n_out is None:
        n_out = W.size(1) if trans_w else W.size(0)
    assert W.dtype == torch.float32 and W.stride(1) == 1
    K = sum(t.size(1) for t, _, _, _ in items)
    if tc is None:
        tc = _PRECISION == "bf16" and M > 0 and _tc_shapes_ok(items, M, n_out, K) and \
            (out_mask is None or _al16(out_mask))
    # 1e-4 mode: split-bf16 tensor-core tiles for all-fp32 layers whose shapes fit
    split = (not tc and _PRECISION == "fp32" and FEATURES["split_tc"] and M > 0 and _tc_shapes_ok(items, M, n_out, K)
             and all(t.dtype == torch.float32 for t, _, _, _ in items) and (out_mask is None or _al16(out_mask))
             and mask_bits is None and bits_out is None and (out is None or out.dtype == torch.float32)
             and all(t.dtype == torch.float32 and _al16(t) for t, _ in (adds or [])))
    if out is None:
        out = torch.empty((M, n_out), dtype=out_dtype if tc else torch.float32, device=W.device)
    if bias is not None:
        assert bias.is_contiguous() and bias.numel() == n_out
    add_segs, nadd = None, 0
    if adds:
        add_segs, nadd = L.make_segs([(t, i, None, 0) for t, i in adds]), len(adds)
        assert all(t.size(1) == n_out for t, _ in adds)
        assert tc or all(t.dtype == torch.float32 for t, _ in adds)
    if mask_bits is not None:
        assert tc and out_mask is None and mask_bits.shape == ((n_out + 31) // 32, M) and mask_bits.is_contiguous()
        m_ptr, m_ld, m_dt = L.ptr(mask_bits), 0, L.BITS
    else:
        m_ptr, m_ld = L.ptr(out_mask), out_mask.stride(0) if out_mask is not None else 0
        m_dt = _DT[out_mask.dtype] if out_mask is not None else 0
    if bits_out is not None:
        assert tc and bits_out.shape == ((n_out + 31) // 32, M) and bits_out.is_contiguous()
    if tc and _tma_ok(items, K, accumulate):
        wr = _packed_weight_segs(W, trans_w, [t.size(1) for t, _, _, _ in items]) if _tma_needs_seg_pack(items) \
            else _packed_weight(W, trans_w, rowmajor=True)
        L.check(L.lib().b3d_linear_tma(segs, len(items), L.ptr(wr), n_out, K, L.ptr(bias), L.ptr(out),
                                       out.stride(0), _DT[out.dtype], M, act, 0, m_ptr, m_ld, m_dt, L.ptr(row_mask),
                                       add_segs, nadd, L.ptr(bits_out), L.stream()), "b3d_linear_tma")
        return out
    if tc or split:
        wp = _packed_weight(W, trans_w, split=split)
        L.check(L.lib().b3d_linear_tc(segs, len(items), L.ptr(wp), n_out, K, L.ptr(bias), L.ptr(out),
                                      out.stride(0), _DT[out.dtype], M, act,
                                      (L.FLAG_ACCUMULATE if accumulate else 0) | (L.FLAG_SPLIT if split else 0),
                                      m_ptr, m_ld, m_dt, L.ptr(row_mask), add_segs, nadd, L.ptr(bits_out),
                                      L.stream()), "b3d_linear_tc")
        return out
    assert out.dtype == torch.float32 and (out_mask is None or out_mask.dtype == torch.float32)
    L.check(L.lib().b3d_linear(segs, len(items), L.ptr(W), W.stride(0), int(trans_w), L.ptr(bias),
                               L.ptr(out), out.stride(0), M, n_out, act,
                               L.FLAG_ACCUMULATE if accumulate else 0, L.ptr(out_mask),
                               out_mask.stride(0) if out_mask is not None else 0, L.ptr(row_mask),
                               add_segs, nadd, L.stream()), "b3d_linear")
    return out


def linear_raw(items, W, bias, M, act=L.ACT_NONE, trans_w=False, out=None, accumulate=False, **kw):
    if _OPTIME is None:
        return _linear_raw_impl(items, W, bias, M, act, trans_w, out, accumulate, **kw)
    K = sum(t.size(1) for t, _, _, _ in items)
    n_out = kw.get("n_out") or (W.size(1) if trans_w else W.size(0))
    es = lambda t: 2 if t.dtype == torch.bfloat16 else 4
    od = kw.get("out_dtype", torch.float32)
    nb = sum(M * t.size(1) * es(t) for t, _, _, _ in items) + M * n_out * (2 if od == torch.bfloat16 else 4)
    if kw.get("out_mask") is not None:
        nb += M * n_out * es(kw["out_mask"])
    if kw.get("mask_bits") is not None:
        nb += M * n_out // 8
    if kw.get("bits_out") is not None:
        nb += M * n_out // 8
    for t, _ in (kw.get("adds") or []):
        nb += M * n_out * es(t)
    sig = ("dgrad" if trans_w else "fwd", M, K, n_out, tuple(str(t.dtype)[6:] + ("g" if i is not None else "") for t, i, _, _ in items),
           str(od)[6:], "mask" if kw.get("out_mask") is not None else ("bits" if kw.get("mask_bits") is not None else ""),
           len(kw.get("adds") or []))
    with _timed(sig, nb):
        return _linear_raw_impl(items, W, bias, M, act, trans_w, out, accumulate, **kw)


def wgrad_raw(dy_item, items, M, n_out, K, dW=None, db=None, accumulate=False, want_bias=True, tc=None):
    if _OPTIME is not None:
        es = lambda t: 2 if t.dtype == torch.bfloat16 else 4
        nb = sum(M * t.size(1) * es(t) for t, _, _, _ in items) + M * n_out * es(dy_item[0])
        sig = ("wgrad", M, K, n_out, tuple(str(t.dtype)[6:] + ("g" if i is not None else "") for t, i, _, _ in items),
               str(dy_item[0].dtype)[6:], "", 0)
        with _timed(sig, nb):
            return _wgrad_raw_impl(dy_item, items, M, n_out, K, dW, db, accumulate, want_bias, tc)
    return _wgrad_raw_impl(dy_item, items, M, n_out, K, dW, db, accumulate, want_bias, tc)


def _wgrad_raw_impl(dy_item, items, M, n_out, K, dW=None, db=None, accumulate=False, want_bias=True, tc=None):
    dev = dy_item[0].device
    if dW is None:
        dW = torch.empty((n_out, K), dtype=torch.float32, device=dev)
    if db is None and want_bias:
        db = torch.empty(n_out, dtype=torch.float32, device=dev)
    lib = L.lib()
    if tc is None:
        tc = _PRECISION == "bf16" and M > 0 and _tc_shapes_ok(items, M, max(n_out, 16), K) and \
            _al16(dy_item[0]) and dy_item[2] is None
    if tc and dy_item[0].dtype == torch.bfloat16 and n_out % 8 == 0 and _tma_dense_ok(items, K, False):
        wsb = lib.b3d_wgrad_tma_workspace_bytes(M, n_out, K)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        L.check(lib.b3d_wgrad_tma(L.make_segs([dy_item]), L.make_segs(items), len(items), L.ptr(dW), dW.stride(0),
                                  L.ptr(db), M, n_out, L.FLAG_ACCUMULATE if accumulate else 0, L.ptr(ws), wsb,
                                  L.stream()), "b3d_wgrad_tma")
        return dW, db
    split = (not tc and _PRECISION == "fp32" and FEATURES["split_tc"] and M > 0 and _tc_shapes_ok(items, M, max(n_out, 16), K)
             and _al16(dy_item[0]) and dy_item[2] is None and dy_item[0].dtype == torch.float32
             and all(t.dtype == torch.float32 for t, _, _, _ in items))
    if tc or split:
        wsb = lib.b3d_wgrad_tc_workspace_bytes(M, n_out, K)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        L.check(lib.b3d_wgrad_tc(L.make_segs([dy_item]), L.make_segs(items), len(items), L.ptr(dW), dW.stride(0),
                                 L.ptr(db), M, n_out, (L.FLAG_ACCUMULATE if accumulate else 0) | (L.FLAG_SPLIT if split else 0),
                                 L.ptr(ws), wsb, L.stream()), "b3d_wgrad_tc")
        return dW, db
    wsb = lib.b3d_wgrad_workspace_bytes(M, n_out, K)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    L.check(lib.b3d_wgrad(L.make_segs([dy_item]), L.make_segs(items), len(items), L.ptr(dW), dW.stride(0),
                          L.ptr(db), M, n_out, L.FLAG_ACCUMULATE if accumulate else 0, L.ptr(ws), wsb,
                          L.stream()), "b3d_wgrad")
    return dW, db


def segment_sum_raw(src, nidx, out=None, accumulate=False, out_dtype=torch.float32):
    src = _rows(src)
    C_ = src.size(1)
    if src.dtype == torch.bfloat16 and (C_ % 8 or not _al16(src)):
        src = src.float()
    if src.dtype != torch.bfloat16 or accumulate:
        out_dtype = torch.float32          # bf16 sums are written by the bf16-source kernel only
    if out is None:
        out = torch.empty((nidx.n, C_), dtype=out_dtype, device=src.device)
    perm = None if nidx.sorted else nidx.perm
    flags = (L.FLAG_ACCUMULATE if accumulate else 0) | (L.FLAG_OUT_BF16 if out.dtype == torch.bfloat16 else 0)
    L.check(L.lib().b3d_segment_sum(L.ptr(src), _DT[src.dtype], src.stride(0), L.ptr(perm), L.ptr(nidx.rowptr), nidx.n, C_,
                                    L.ptr(out), out.stride(0), flags, L.stream()), "b3d_segment_sum")
    return out


def gather_rows_raw(src, idx32, out_dtype=torch.float32, relu_mask=None, relu_bits=None):
    src = _rows(src)
    if src.dtype == torch.bfloat16 and not (out_dtype == torch.bfloat16 and src.size(1) % 8 == 0 and _al16(src)):
        src = src.float()
    M = idx32.numel()
    if out_dtype == torch.bfloat16 and (src.size(1) % 8 or not _al16(src)):
        out_dtype = torch.float32
    out = torch.empty((M, src.size(1)), dtype=out_dtype, device=src.device)
    fused_mask = relu_mask if (relu_mask is not None and out_dtype == torch.bfloat16 and
                               relu_mask.dtype == torch.bfloat16 and _al16(relu_mask)) else None
    if relu_bits is not None and out_dtype == torch.bfloat16:      # sign bits: 1/16 of the mask bytes
        L.check(L.lib().b3d_gather_rows(L.ptr(src), src.stride(0), L.ptr(idx32), M, src.size(1), L.ptr(out),
                                        _DT[out_dtype], out.stride(0), L.ptr(relu_bits), 0, L.BITS, _DT[src.dtype],
                                        L.stream()),
                "b3d_gather_rows")
        return out
    mdt = L.BF16
    if fused_mask is None and relu_mask is not None and out_dtype == torch.float32 and relu_mask.dtype == torch.float32 \
            and relu_mask.dim() == 2 and relu_mask.stride(1) == 1 and relu_mask.shape == out.shape:
        fused_mask, mdt = relu_mask, L.F32          # fp32 gather + fp32 ReLU mask in one pass (1e-4 modes)
    L.check(L.lib().b3d_gather_rows(L.ptr(src), src.stride(0), L.ptr(idx32), M, src.size(1), L.ptr(out),
                                    _DT[out_dtype], out.stride(0), L.ptr(fused_mask),
                                    fused_mask.stride(0) if fused_mask is not None else 0, mdt, _DT[src.dtype],
                                    L.stream()),
            "b3d_gather_rows")
    if relu_mask is not None and fused_mask is None:
        out = out * (relu_mask > 0)
    return out


def row_nonzero(feats):
    """bool [N]: sum(feats[n]) != 0 — the modality-present predicate (clr_att_gnn.py:111-121)."""
    f = feats.reshape(feats.size(0), -1).contiguous().float()
    mask = torch.empty(f.size(0), dtype=torch.uint8, device=f.device)
    L.check(L.lib().b3d_row_nonzero(L.ptr(f), f.size(1), f.size(0), L.ptr(mask), L.stream()), "b3d_row_nonzero")
    return mask.bool()


def knn_frames(x, frame_ptr, k):
    """[N,k] int64 neighbour table (global ids, -1 padded), SURVEY A.5 order."""
    x = _rows(x.detach().float())
    N, D = x.shape
    fp = frame_ptr.to(device=x.device, dtype=torch.int32).contiguous()
    F = fp.numel() - 1
    out = torch.empty((N, k), dtype=torch.int64, device=x.device)
    scratch = torch.empty(F + 1, dtype=torch.int32, device=x.device)
    L.check(L.lib().b3d_knn_frames(L.ptr(x), x.stride(0), D, L.ptr(fp), F, N, k, L.ptr(out), L.ptr(scratch),
                                   L.stream()), "b3d_knn_frames")
    return out


def _gat_forward(h, att_src, att_dst, bias, nbr, slope, want_alpha):
    h = _rows(h.float())
    N, D = h.shape
    out = torch.empty((N, D), dtype=torch.float32, device=h.device)
    scratch = torch.empty(2 * N, dtype=torch.float32, device=h.device)
    alpha = torch.empty((N, nbr.size(1)), dtype=torch.float32, device=h.device) if want_alpha else None
    L.check(L.lib().b3d_gat_aggregate(L.ptr(h), h.stride(0), D, L.ptr(att_src.reshape(-1).contiguous()),
                                      L.ptr(att_dst.reshape(-1).contiguous()), L.ptr(bias), L.ptr(nbr),
                                      nbr.size(1), N, slope, L.ptr(out), out.stride(0), L.ptr(alpha), L.ptr(scratch),
                                      L.stream()), "b3d_gat_aggregate")
    return out, alpha, scratch, h


class _GATAggregate(torch.autograd.Function):
    """GATConv aggregation over a padded neighbour table with its backward (b3d_gat_bwd): softmax /
    LeakyReLU / score paths per target, then a deterministic per-source reduction over the reversed
    table (CSR built by the device radix sort, padding mapped to a dummy node)."""

    @staticmethod
    def forward(ctx, h, att_src, att_dst, bias, nbr, slope):
        out, alpha, scratch, hf = _gat_forward(h, att_src, att_dst, bias, nbr, slope, True)
        ctx.slope, ctx.shapes = slope, (att_src.shape, att_dst.shape, h.dtype)
        ctx.save_for_backward(hf, att_src.reshape(-1).contiguous(), att_dst.reshape(-1).contiguous(), nbr, alpha, scratch)
        ctx.has_bias = bias is not None
        return out

    @staticmethod
    def backward(ctx, dout):
        h, a_src, a_dst, nbr, alpha, a_s_a_d = ctx.saved_tensors
        N, D = h.shape
        k = nbr.size(1)
        g = _rows(dout.float())
        if not g.is_contiguous():
            g = g.contiguous()
        # reversed table: entry p = t*k + l is an edge (source nbr[t,l] or dummy N) -> target t
        src = torch.where(nbr < 0, torch.full_like(nbr, N), nbr).reshape(-1)
        tgt = torch.arange(N * k, device=nbr.device, dtype=torch.int64) // k
        rev = Graph(torch.stack([src, tgt]), N + 1)
        dh = torch.empty((N, D), dtype=torch.float32, device=h.device)
        ds = torch.empty((N, k), dtype=torch.float32, device=h.device)
        das_dad = torch.empty(2 * N, dtype=torch.float32, device=h.device)
        L.check(L.lib().b3d_gat_bwd(L.ptr(g), g.stride(0), L.ptr(h), h.stride(0), D, L.ptr(a_src), L.ptr(a_dst),
                                    L.ptr(nbr), k, N, ctx.slope, L.ptr(alpha), L.ptr(a_s_a_d), L.ptr(rev.rowptr_src),
                                    L.ptr(rev.perm_src), L.ptr(dh), dh.stride(0), L.ptr(ds), L.ptr(das_dad),
                                    L.stream()), "b3d_gat_bwd")
        # d att_src = das^T h, d att_dst = dad^T h, d bias = colsum(dout): small deterministic wgrads
        S = das_dad.view(2, N).t().contiguous()
        datt, _ = wgrad_raw((S, None, None, 0), [(h, None, None, 0)], N, 2, D, want_bias=False, tc=False)
        _, dbias = wgrad_raw((g, None, None, 0), [(S, None, None, 0)], N, D, 2, want_bias=True, tc=False)
        s_shape, d_shape, h_dtype = ctx.shapes
        return (dh.to(h_dtype), datt[0].reshape(s_shape), datt[1].reshape(d_shape),
                dbias if ctx.has_bias else None, None, None)


def gat_aggregate(h, att_src, att_dst, bias, nbr, slope=0.2):
    """GATConv(heads=1, add_self_loops=False) aggregation; differentiable w.r.t. h, att_src, att_dst, bias."""
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in (h, att_src, att_dst, bias)):
        return _GATAggregate.apply(h, att_src, att_dst, bias, nbr, slope)
    return _gat_forward(h, att_src, att_dst, bias, nbr, slope, False)[0]


def adam_step(p, g, m, v, lr, betas, eps, weight_decay, step, grad_scale=1.0):
    """In-place Adam on flat fp32 buffers (train.py:106-109)."""
    L.check(L.lib().b3d_adam_step(L.ptr(p), L.ptr(g), L.ptr(m), L.ptr(v), p.numel(), lr, betas[0], betas[1],
                                  eps, weight_decay, step, grad_scale, L.stream()), "b3d_adam_step")


def adam_step_dev(p, g, m, v, lr, betas, eps, weight_decay, step_counter, grad_scale=1.0):
    """adam_step with the step number read from (and advanced in) the int32 device tensor `step_counter`."""
    L.check(L.lib().b3d_adam_step_dev(L.ptr(p), L.ptr(g), L.ptr(m), L.ptr(v), p.numel(), lr, betas[0], betas[1],
                                      eps, weight_decay, L.ptr(step_counter), grad_scale, L.stream()), "b3d_adam_step_dev")


# ----------------------------------------------------------------------------- autograd
_ACT = {None: L.ACT_NONE, "relu": L.ACT_RELU, "sigmoid": L.ACT_SIGMOID}
_MASK = {None: L.MASK_NONE, "relu": L.MASK_RELU, "sigmoid": L.MASK_SIGMOID}


class _FusedMLP(torch.autograd.Function):
    """A whole nn.Sequential(Linear, ReLU, ..., Linear[, final act]) over a (virtually) concatenated,
    optionally gathered input: y = MLP(cat_s(gather(x_s, idx_s))) [rows zeroed by row_mask].

    Forward: one libb3d GEMM per layer with fused gather/concat/bias/activation. In bf16 mode the
    hidden activations are stored in bf16 (they are rounded to bf16 by the next layer's tensor-core
    tile anyway); the chain output stays fp32.
    Backward, layer by layer: dW/db by the deterministic split-row wgrad; the input gradient of
    layer l is produced by the transposed GEMM whose epilogue applies the ReLU mask of layer l-1
    (so it IS the pre-activation gradient layer l-1 needs, no extra pass); gathered input segments
    are reduced back to nodes with the CSR segmented sum (no atomics)."""

    @staticmethod
    def forward(ctx, nl, final_act, row_mask, nidx, add_nidx, out_dtype, premasked, *tensors):
        Ws = [w if w.stride(1) == 1 else w.contiguous() for w in tensors[0:2 * nl:2]]
        bs = [b.contiguous() if b is not None else None for b in tensors[1:2 * nl:2]]
        nx = len(nidx)
        xs = [_rows(x) for x in tensors[2 * nl:2 * nl + nx]]
        add_ts = [_rows(t) for t in tensors[2 * nl + nx:]]
        assert len(add_ts) == len(add_nidx) and (not add_ts or (final_act is None or nl > 1 or premasked))
        assert not premasked or final_act == "relu"
        ctx.premasked = premasked
        W0dev = Ws[0].device
        ctx.add_dtypes = [t.dtype for t in add_ts]
        adds = [(t, ni.idx if ni is not None else None) for t, ni in zip(add_ts, add_nidx)]
        M = nidx[0].idx.numel() if nidx[0] is not None else xs[0].size(0)
        for x, ni in zip(xs, nidx):
            assert (ni.idx.numel() if ni is not None else x.size(0)) == M, "segment row counts differ"
        assert sum(x.size(1) for x in xs) == Ws[0].size(1), "concatenated width != weight in_features"
        items = [(x, ni.idx if ni is not None else None, None, 0) for x, ni in zip(xs, nidx)]
        rm = row_mask.to(torch.uint8).contiguous() if row_mask is not None else None
        # the whole chain runs on tensor cores or not at all (keeps dtypes of saved activations uniform)
        tc = _PRECISION == "bf16" and M > 0 and (final_act is None or premasked) and \
            _tc_shapes_ok(items, M, Ws[0].size(0), Ws[0].size(1)) \
            and all(w.size(0) % 8 == 0 and w.size(1) % 8 == 0 and w.size(0) >= 16 and w.size(1) >= 32 for w in Ws)
        if not tc:
            adds = [(t.float(), i) for t, i in adds]
        if tc and FEATURES["bf16_inputs"] and any(x.dtype == torch.float32 and ni is None for x, ni in zip(xs, nidx)):
            # dense fp32 inputs of a bf16 chain (the raw sensor features of the fc_*_encoder layers): rounded to bf16
            # ONCE here — exactly what the thread-staged kernel does per tile while staging — so that the layer and its
            # weight gradient run on the TMA-fed kernels (k_wgrad_tc on fp32 rows ran at ~1 TB/s)
            xs = [x.to(torch.bfloat16) if (x.dtype == torch.float32 and ni is None and x.size(1) % 8 == 0) else x
                  for x, ni in zip(xs, nidx)]
            items = [(x, ni.idx if ni is not None else None, None, 0) for x, ni in zip(xs, nidx)]
        if not tc and any(x.dtype != torch.float32 for x in xs):      # fp32 kernels take fp32 operands
            xs = [x.float() for x in xs]
            items = [(x, ni.idx if ni is not None else None, None, 0) for x, ni in zip(xs, nidx)]
        ctx.in_dtypes = [t.dtype for t in tensors[2 * nl:2 * nl + nx]]
        # chains that are not bf16 end to end keep fp32 activations; each launch then picks the bf16
        # tile kernel on its own when the precision mode allows and its shapes fit (tc=None = auto)
        tc_arg = True if tc else None
        # bf16 chains record every ReLU output's SIGN BITS (1 bit per element, written by the producing
        # epilogue): the backward pass masks gradients from them instead of re-reading the activation
        want_bits = any(ctx.needs_input_grad)      # inference (no_grad): nothing reads them
        acts, bits, cur = [], [], items
        # the whole chain as ONE launch: hidden activations stay in shared memory (chain_tc.cu). Training keeps
        # the same saved tensors as the per-layer path (bf16 activations + sign bits), inference writes the output only
        ctx.chain = False
        if (tc and _USE_CHAIN and nl >= 2 and nl <= L.CHAIN_MAX_LAYERS and rm is None and _tma_dense_ok(items, Ws[0].size(1), False)
                and (out_dtype or torch.float32) == torch.bfloat16 and _chain_dims_ok([w.size(0) for w in Ws])
                and Ws[0].size(1) % 16 == 0 and len(adds) <= 2 and all(t.dtype == torch.bfloat16 and _al16(t) for t, _ in adds)):
            idxs = []
            for _, i in adds:
                if i is not None and not any(i is j for j in idxs):
                    idxs.append(i)
            if len(idxs) <= 2:
                specs = []
                for l in range(nl):
                    last = l == nl - 1
                    relu_out = (not last) or premasked
                    nb = Ws[l].size(0)
                    y = torch.empty((M, nb), dtype=torch.bfloat16, device=W0dev) if (last or want_bits) else None
                    bt = new_relu_bits(M, nb, W0dev) if (relu_out and _USE_BITS and want_bits) else None
                    specs.append(dict(W=Ws[l], bias=bs[l], act=L.ACT_RELU if relu_out else L.ACT_NONE, out=y, bits_out=bt,
                                      adds=[(t, -1 if i is None else [k for k, j in enumerate(idxs) if j is i][0])
                                            for t, i in adds] if l == 0 else None))
                    acts.append(y)
                    bits.append(bt)
                ctx.chain = chain_run([t for t, _, _, _ in items], specs, idxs[0] if idxs else None,
                                      idxs[1] if len(idxs) > 1 else None, M)
                if not ctx.chain:
                    acts, bits = [], []
        for l in range(nl if not ctx.chain else 0):
            last = l == nl - 1
            relu_out = (not last) or premasked
            nb = Ws[l].size(0)
            bt = new_relu_bits(M, nb, W0dev) if (tc and relu_out and _USE_BITS and nb % 32 == 0 and want_bits) else None
            y = linear_raw(cur, Ws[l], bs[l], M, _ACT[final_act] if last else L.ACT_RELU,
                           row_mask=rm if last else None, tc=tc_arg,
                           out_dtype=((out_dtype or torch.float32) if last else torch.bfloat16) if tc else torch.float32,
                           adds=adds if l == 0 else None, bits_out=bt)
            acts.append(y)
            bits.append(bt)
            cur = [(y, None, None, 0)]
        ctx.nl, ctx.final_act, ctx.rm, ctx.nidx, ctx.M, ctx.tc = nl, final_act, rm, nidx, M, tc
        ctx.add_nidx = add_nidx
        ctx.has_bias = [b is not None for b in bs]
        ctx.bits = bits
        ctx.save_for_backward(*Ws, *[t if t is not None else acts[-1] for t in acts], *xs, *[b for b in bs if b is not None])
        if premasked:      # the consumer (segment_sum) needs the output's sign bits for its own backward
            out_bits = bits[-1] if bits[-1] is not None else acts[-1].new_zeros(0, dtype=torch.int32)
            ctx.mark_non_differentiable(out_bits)
            return acts[-1], out_bits
        return acts[-1]

    @staticmethod
    def backward(ctx, dy, *_unused):
        nl, nidx, M, tc = ctx.nl, ctx.nidx, ctx.M, ctx.tc
        tc_arg = True if tc else None
        saved = ctx.saved_tensors
        Ws, acts, xs = saved[:nl], saved[nl:2 * nl], saved[2 * nl:]
        dz = _rows(dy)
        if dz.dtype != torch.float32 and not tc:
            dz = dz.float()
        if ctx.rm is not None:
            dz = dz * ctx.rm.unsqueeze(1)
        # only a chain-final activation needs an operand mask (fp32 path; tc chains end linear)
        # (premasked: the consumer's backward already applied the final ReLU's mask to dy)
        need_mask = ctx.final_act is not None and not ctx.premasked
        dz_item = (dz, None, acts[-1] if need_mask else None, _MASK[ctx.final_act] if need_mask else 0)
        items0 = [(x, ni.idx if ni is not None else None, None, 0) for x, ni in zip(xs, nidx)]
        grads = [None] * (2 * nl)
        nx = len(nidx)
        need_x = ctx.needs_input_grad[7 + 2 * nl:7 + 2 * nl + nx]
        need_add = ctx.needs_input_grad[7 + 2 * nl + nx:]
        b_saved = list(xs[nx:])
        bs = [b_saved.pop(0) if hb else None for hb in ctx.has_bias]
        xs = xs[:nx]
        dxs = [None] * nx
        dadds = [None] * len(ctx.add_nidx)
        # all input-gradient GEMMs of the chain as ONE fused launch (dZ_{l-1} = (dZ_l W_l) * relu'(z_{l-1}), the
        # gradient tiles stay in shared memory between layers; every dZ_l is also written for its weight gradient)
        pre_dz, pre_dA = None, None
        low_x = tc and all(dt == torch.bfloat16 for dt, nd in zip(ctx.in_dtypes, need_x) if nd)
        if (getattr(ctx, "chain", False) and not need_mask and nl >= 2 and all(b is not None for b in ctx.bits[:nl - 1])
                and Ws[-1].size(0) % 16 == 0):
            dzc = dz if (dz.dtype == torch.bfloat16 and _al16(dz)) else dz.to(torch.bfloat16).contiguous()
            specs = [dict(W=Ws[l], transpose=True, act=L.ACT_MASKBITS, bits_in=ctx.bits[l - 1],
                          out=torch.empty((M, Ws[l].size(1)), dtype=torch.bfloat16, device=dz.device))
                     for l in range(nl - 1, 0, -1)]
            with_x = any(need_x) and low_x and Ws[0].size(1) % 64 == 0 and Ws[0].size(1) <= 512
            if with_x:
                specs.append(dict(W=Ws[0], transpose=True, act=L.ACT_NONE,
                                  out=torch.empty((M, Ws[0].size(1)), dtype=torch.bfloat16, device=dz.device)))
            if len(specs) <= L.CHAIN_MAX_LAYERS and chain_run([dzc], specs, None, None, M):
                pre_dz = {l - 1: sp["out"] for l, sp in zip(range(nl - 1, 0, -1), specs)}
                pre_dA = specs[-1]["out"] if with_x else None
                dz_item = (dzc, None, None, 0)
        for l in range(nl - 1, -1, -1):
            W = Ws[l]
            n_out, K = W.shape
            a_items = items0 if l == 0 else [(acts[l - 1], None, None, 0)]
            if l == 0:
                # dz is now the gradient of layer 0's pre-activation: the node-side addends receive
                # its per-node sums (deterministic CSR reduction; identity index = plain copy)
                for t, ni in enumerate(ctx.add_nidx):
                    if need_add[t]:
                        assert dz_item[2] is None
                        g = segment_sum_raw(dz_item[0], ni, out_dtype=ctx.add_dtypes[t]) if ni is not None else dz_item[0]
                        dadds[t] = g if g.dtype == ctx.add_dtypes[t] else g.to(ctx.add_dtypes[t])
            if ctx.needs_input_grad[7 + 2 * l]:
                tW = _grad_target(W)
                tb = _grad_target(bs[l]) if ctx.has_bias[l] else None
                if tW is not None and (not ctx.has_bias[l] or tb is not None):
                    # grads_into_params: accumulated into the parameters' .grad by the kernel, nothing for autograd
                    wgrad_raw(dz_item, a_items, M, n_out, K, dW=tW, db=tb, accumulate=True, want_bias=ctx.has_bias[l], tc=tc_arg)
                else:
                    dW, db = wgrad_raw(dz_item, a_items, M, n_out, K, want_bias=ctx.has_bias[l], tc=tc_arg)
                    grads[2 * l], grads[2 * l + 1] = dW, (db if ctx.has_bias[l] else None)
            if l > 0:
                if pre_dz is not None:
                    dz = pre_dz[l - 1]
                else:
                    mb = ctx.bits[l - 1]
                    dz = linear_raw([dz_item], W, None, M, trans_w=True, out_mask=acts[l - 1] if mb is None else None,
                                    mask_bits=mb, tc=tc_arg, out_dtype=torch.bfloat16 if tc else torch.float32)
                dz_item = (dz, None, None, 0)
            elif any(need_x):
                # [M, K]; bf16 when every consumer of the slices is a bf16 tensor (halves the largest
                # backward tensor), fp32 otherwise
                dA = pre_dA if pre_dA is not None else \
                    linear_raw([dz_item], W, None, M, trans_w=True, tc=tc_arg,
                               out_dtype=torch.bfloat16 if low_x else torch.float32)
                off = 0
                for s, (x, ni) in enumerate(zip(xs, nidx)):
                    w = x.size(1)
                    if need_x[s]:
                        sl = dA[:, off:off + w]
                        g = segment_sum_raw(sl, ni) if ni is not None else sl
                        dxs[s] = g if g.dtype == ctx.in_dtypes[s] else g.to(ctx.in_dtypes[s])
                    off += w
        return (None, None, None, None, None, None, None, *grads, *dxs, *dadds)


class _MPEdgeBlock(torch.autograd.Function):
    """The edge side of one message-passing iteration as ONE fused launch per direction (chain_tc.cu):
        e'  = edge_update(cat[x_i, x_j, e, att])                         (clr_att_gnn.py:314-317)
        h_f = relu(W_f . cat[x_i, e', x0_i]),  h_p = relu(W_p . cat[x_j, e', x0_j])   (first layers, :319-327)
    in the pre-projected form (node-side weight blocks applied per node: p_i / p_j / p_f / p_p arrive as
    row-gathered bf16 addends). Forward program: 128 -> 256 -> 128 -> 64 (= e') -> {192, 192}; the hidden tiles and
    e' feed the next layer from shared memory. Backward program: cat[dh_f, dh_p] -> de' (+ the direct gradient of e')
    -> 128 -> 256 -> 128, every tile masked by the forward sign bits; the five weight gradients and the four
    per-node addend gradients (deterministic CSR sums) follow from the written gradient tiles."""

    @staticmethod
    def forward(ctx, g, e, att, p_i, p_j, p_f, p_p, W0, W1, b1, W2, b2, Wf, Wp):
        M = e.size(0)
        dev = e.device
        train = any(ctx.needs_input_grad)
        bf = torch.bfloat16
        n0, n1, n2, nm = W0.size(0), W1.size(0), W2.size(0), Wf.size(0)
        mk = lambda n: torch.empty((M, n), dtype=bf, device=dev)
        x1, x2 = (mk(n0), mk(n1)) if train else (None, None)
        e_new, h_f, h_p = mk(n2), mk(nm), mk(nm)
        bt = lambda n: new_relu_bits(M, n, dev) if train else None
        b_x1, b_x2, b_f, b_p = bt(n0), bt(n1), bt(nm), bt(nm)
        specs = [dict(W=W0, src=-1, act=L.ACT_RELU, adds=[(p_i, 0), (p_j, 1)], out=x1, bits_out=b_x1),
                 dict(W=W1, src=0, act=L.ACT_RELU, bias=b1, out=x2, bits_out=b_x2),
                 dict(W=W2, src=1, act=L.ACT_NONE, bias=b2, out=e_new),
                 dict(W=Wf, src=2, act=L.ACT_RELU, adds=[(p_f, 0)], out=h_f, bits_out=b_f),
                 dict(W=Wp, src=2, act=L.ACT_RELU, adds=[(p_p, 1)], out=h_p, bits_out=b_p)]
        ins = [e] + ([att] if att is not None else [])
        if not chain_run(ins, specs, g.by_dst.idx, g.by_src.idx, M):
            raise RuntimeError("mp_edge_block: chain plan rejected (call mp_edge_block_supported first)")
        ctx.g, ctx.has_att = g, att is not None
        if train:
            ctx.save_for_backward(e, att if att is not None else e, x1, x2, e_new, b_x1, b_x2, W0, W1, W2, Wf, Wp)
        empty = torch.zeros(0, dtype=torch.int32, device=dev)
        outs = (e_new, h_f, b_f if train else empty, h_p, b_p if train else empty)
        ctx.mark_non_differentiable(outs[2], outs[4])
        return outs

    @staticmethod
    def backward(ctx, de, dhf, _bf, dhp, _bp):
        e, att, x1, x2, e_new, b_x1, b_x2, W0, W1, W2, Wf, Wp = ctx.saved_tensors
        g = ctx.g
        M, dev, bf = e.size(0), e.device, torch.bfloat16
        n0, n1, n2, nm, k0 = W0.size(0), W1.size(0), W2.size(0), Wf.size(0), W0.size(1)
        z = lambda t, n: t.contiguous() if t is not None else torch.zeros((M, n), dtype=bf, device=dev)
        dhf, dhp = z(dhf, nm), z(dhp, nm)
        mk = lambda n: torch.empty((M, n), dtype=bf, device=dev)
        dz2, dz1, dz0, dA = mk(n2), mk(n1), mk(n0), mk(k0)
        w_fp = torch.cat([Wf, Wp], 0)                      # [2*nm, n2]: de' = cat[dh_f, dh_p] . [W_f ; W_p] (+ de)
        specs = [dict(W=w_fp, transpose=True, src=-1, act=L.ACT_NONE, out=dz2,
                      adds=[(de.contiguous(), -1)] if de is not None else None),
                 dict(W=W2, transpose=True, src=0, act=L.ACT_MASKBITS, bits_in=b_x2, out=dz1),
                 dict(W=W1, transpose=True, src=1, act=L.ACT_MASKBITS, bits_in=b_x1, out=dz0),
                 dict(W=W0, transpose=True, src=2, act=L.ACT_NONE, out=dA)]
        if not chain_run([dhf, dhp], specs, None, None, M):
            raise RuntimeError("mp_edge_block backward: chain plan rejected")
        it = lambda t: (t, None, None, 0)
        ins = [it(e)] + ([it(att)] if ctx.has_att else [])
        dW0, _ = wgrad_raw(it(dz0), ins, M, n0, k0, want_bias=False, tc=True)
        dW1, db1 = wgrad_raw(it(dz1), [it(x1)], M, n1, n0, tc=True)
        dW2, db2 = wgrad_raw(it(dz2), [it(x2)], M, n2, n1, tc=True)
        dWf, _ = wgrad_raw(it(dhf), [it(e_new)], M, nm, n2, want_bias=False, tc=True)
        dWp, _ = wgrad_raw(it(dhp), [it(e_new)], M, nm, n2, want_bias=False, tc=True)
        dp_i = segment_sum_raw(dz0, g.by_dst, out_dtype=bf)
        dp_j = segment_sum_raw(dz0, g.by_src, out_dtype=bf)
        dp_f = segment_sum_raw(dhf, g.by_dst, out_dtype=bf)
        dp_p = segment_sum_raw(dhp, g.by_src, out_dtype=bf)
        d_e = dA[:, :e.size(1)]
        d_att = dA[:, e.size(1):] if ctx.has_att else None
        return (None, d_e, d_att, dp_i, dp_j, dp_f, dp_p, dW0, dW1, db1, dW2, db2, dWf, dWp)


def _it(t, idx=None):
    return (t, idx, None, 0)


class _MPEdgeBlockG(torch.autograd.Function):
    """The edge side of one message-passing iteration (clr_att_gnn.py:314-327) with the node-feature operands
    row-gathered by the TMA unit (b3d_linear_tma, tile::gather4):
        X1 = relu(W0 . cat[x_i, x_j, e, att] + b0)       — the reference's own un-projected first layer
        X2 = relu(W1 X1 + b1),  e' = W2 X2 + b2
        h_f = relu(Wf . cat[x_i, e', x0_i] + bf),  h_p = relu(Wp . cat[x_j, e', x0_j] + bp)
    so no epilogue waits on gathered addend rows (the forward tiles with addends ran at 1.8-2.2 TB/s, the plain
    ones at 5.5+; profiles/r1_bf16_launch_summary.md). The BACKWARD keeps the pre-projected form: p_i .. p_p (the
    per-node first-layer blocks, computed by the caller's node-level GEMM and unused by this forward) receive the
    per-node sums of the pre-activation gradients, and autograd carries them on into x, x0 and the node-side weight
    blocks — the same gradients as the un-projected function's, at node-level instead of edge-level cost."""

    @staticmethod
    def forward(ctx, g, gathered, x, x0, e, att, p_i, p_j, p_f, p_p, W0, b0, W1, b1, W2, b2, Wf, bf_, Wp, bp_, W0e, Wfe, Wpe):
        M, dev, bf = e.size(0), e.device, torch.bfloat16
        train = any(ctx.needs_input_grad)
        dst, src = g.by_dst.idx, g.by_src.idx
        bt = lambda n: new_relu_bits(M, n, dev) if train else None
        b_x1, b_x2, b_f, b_p = bt(W0.size(0)), bt(W1.size(0)), bt(Wf.size(0)), bt(Wp.size(0))
        dense = [_it(e)] + ([_it(att)] if att is not None else [])
        if gathered:      # node features row-gathered by the TMA unit into the operand tiles (un-projected first layers)
            x1 = linear_raw([_it(x, dst), _it(x, src)] + dense, W0, b0, M, L.ACT_RELU, tc=True, out_dtype=bf, bits_out=b_x1)
        else:             # per-node pre-projections as row-gathered epilogue addends
            x1 = linear_raw(dense, W0e, None, M, L.ACT_RELU, tc=True, out_dtype=bf, bits_out=b_x1,
                            adds=[(p_i, dst), (p_j, src)])
        x2 = linear_raw([_it(x1)], W1, b1, M, L.ACT_RELU, tc=True, out_dtype=bf, bits_out=b_x2)
        e_new = linear_raw([_it(x2)], W2, b2, M, L.ACT_NONE, tc=True, out_dtype=bf)
        if gathered:
            h_f = linear_raw([_it(x, dst), _it(e_new), _it(x0, dst)], Wf, bf_, M, L.ACT_RELU, tc=True, out_dtype=bf, bits_out=b_f)
            h_p = linear_raw([_it(x, src), _it(e_new), _it(x0, src)], Wp, bp_, M, L.ACT_RELU, tc=True, out_dtype=bf, bits_out=b_p)
        else:
            h_f = linear_raw([_it(e_new)], Wfe, None, M, L.ACT_RELU, tc=True, out_dtype=bf, bits_out=b_f, adds=[(p_f, dst)])
            h_p = linear_raw([_it(e_new)], Wpe, None, M, L.ACT_RELU, tc=True, out_dtype=bf, bits_out=b_p, adds=[(p_p, src)])
        ctx.g, ctx.has_att = g, att is not None
        if train:
            ctx.save_for_backward(e, att if att is not None else e, x1, x2, e_new, b_x1, b_x2, W0e, W1, W2, Wfe, Wpe)
        empty = torch.zeros(0, dtype=torch.int32, device=dev)
        outs = (e_new, h_f, b_f if train else empty, h_p, b_p if train else empty)
        ctx.mark_non_differentiable(outs[2], outs[4])
        return outs

    @staticmethod
    def backward(ctx, de, dhf, _bf, dhp, _bp):
        e, att, x1, x2, e_new, b_x1, b_x2, W0e, W1, W2, Wfe, Wpe = ctx.saved_tensors
        g = ctx.g
        M, dev, bf = e.size(0), e.device, torch.bfloat16
        n0, n1, n2, nm, k0 = W0e.size(0), W1.size(0), W2.size(0), Wfe.size(0), W0e.size(1)
        z = lambda t, n: t.contiguous() if t is not None else torch.zeros((M, n), dtype=bf, device=dev)
        dhf, dhp = z(dhf, nm), z(dhp, nm)
        # de' = cat[dh_f, dh_p] . [Wf_e ; Wp_e] + (direct gradient of e'), then the masked chain back to cat[e, att]
        w_fp = torch.cat([Wfe, Wpe], 0)
        # the direct gradient arrives as a column slice of the next block's [E, 128] input gradient: a dense addend
        # with its own row stride, no copy
        if de is not None and not (de.dim() == 2 and de.stride(1) == 1 and de.dtype == bf and _al16(de) and de.stride(0) % 8 == 0):
            de = de.to(bf).contiguous()
        dz2 = linear_raw([_it(dhf), _it(dhp)], w_fp, None, M, trans_w=True, tc=True, out_dtype=bf,
                         adds=[(de, None)] if de is not None else None)
        dz1 = linear_raw([_it(dz2)], W2, None, M, trans_w=True, mask_bits=b_x2, tc=True, out_dtype=bf)
        dz0 = linear_raw([_it(dz1)], W1, None, M, trans_w=True, mask_bits=b_x1, tc=True, out_dtype=bf)
        dA = linear_raw([_it(dz0)], W0e, None, M, trans_w=True, tc=True, out_dtype=bf)
        ins = [_it(e)] + ([_it(att)] if ctx.has_att else [])
        dW0e, _ = wgrad_raw(_it(dz0), ins, M, n0, k0, want_bias=False, tc=True)
        dW1, db1 = wgrad_raw(_it(dz1), [_it(x1)], M, n1, n0, tc=True)
        dW2, db2 = wgrad_raw(_it(dz2), [_it(x2)], M, n2, n1, tc=True)
        dWfe, _ = wgrad_raw(_it(dhf), [_it(e_new)], M, nm, n2, want_bias=False, tc=True)
        dWpe, _ = wgrad_raw(_it(dhp), [_it(e_new)], M, nm, n2, want_bias=False, tc=True)
        dp_i = segment_sum_raw(dz0, g.by_dst, out_dtype=bf)
        dp_j = segment_sum_raw(dz0, g.by_src, out_dtype=bf)
        dp_f = segment_sum_raw(dhf, g.by_dst, out_dtype=bf)
        dp_p = segment_sum_raw(dhp, g.by_src, out_dtype=bf)
        d_e = dA[:, :e.size(1)]
        d_att = dA[:, e.size(1):] if ctx.has_att else None
        # inputs: g, gathered, x, x0, e, att, p_i, p_j, p_f, p_p, W0, b0, W1, b1, W2, b2, Wf, bf_, Wp, bp_, W0e, Wfe, Wpe
        return (None, None, None, None, d_e, d_att, dp_i, dp_j, dp_f, dp_p, None, None, dW1, db1, dW2, db2, None, None, None,
                None, dW0e, dWfe, dWpe)


def mp_edge_block_explicit_supported(x, e, att):
    """The explicit edge block (one autograd node per iteration instead of five chains + two fan-outs): bf16 mode,
    dense 16-byte aligned bf16 edge tensors. Its forward takes the TMA row gathers when FEATURES["gather_tma"] is on,
    the per-node addends otherwise; its backward forms de' in ONE K-concatenated GEMM with the direct gradient as an
    addend (no 3-way sum kernel, one launch instead of two)."""
    if _PRECISION != "bf16" or e.size(0) < _TC_MIN_ROWS or not FEATURES["edge_block"]:
        return False
    ts = [x, e] + ([att] if att is not None else [])
    return all(t.dtype == torch.bfloat16 and _al16(t) and t.size(1) % 8 == 0 and t.is_contiguous() for t in ts) \
        and e.size(1) % 64 == 0 and (att is None or att.size(1) % 64 == 0)


def mp_edge_block_explicit(g, x, x0, e, att, p_i, p_j, p_f, p_p, eu, lf0, lp0, D, E_):
    """eu: the three nn.Linear of edge_update; lf0 / lp0: first layers of create_future_msgs / create_past_msgs.
    Returns (e', h_f, bits_f, h_p, bits_p). The full first-layer weights and biases feed the (gathered) forward only
    (this block returns no gradient for them: the node-side blocks and biases get theirs through p_i .. p_p); the
    edge-feature column blocks are passed again as views so that their gradients come back from this block."""
    return _MPEdgeBlockG.apply(g, FEATURES["gather_tma"], x, x0, e, att, p_i, p_j, p_f, p_p, eu[0].weight, eu[0].bias,
                               eu[1].weight, eu[1].bias, eu[2].weight, eu[2].bias, lf0.weight, lf0.bias,
                               lp0.weight, lp0.bias, eu[0].weight[:, 2 * D:], lf0.weight[:, D:D + E_],
                               lp0.weight[:, D:D + E_])


def mp_edge_block_supported(e, att, W0, W1, W2, Wf):
    """The fused edge block needs dense bf16 edge tensors and widths the chain kernel plans (multimodal model:
    64|64 -> 256 -> 128 -> 64 -> 192; the poses-only widths 96 / 32 stay on the per-layer kernels)."""
    if not _USE_CHAIN or _PRECISION != "bf16" or e.size(0) < _TC_MIN_ROWS:
        return False
    ins = [e] + ([att] if att is not None else [])
    if any(t.dtype != torch.bfloat16 or not _al16(t) or t.size(1) % 64 for t in ins):
        return False
    return _chain_dims_ok([W0.size(0), W1.size(0), W2.size(0), Wf.size(0), W0.size(1)]) and 2 * Wf.size(0) <= 512


def mp_edge_block(g, e, att, p_i, p_j, p_f, p_p, W0, W1, b1, W2, b2, Wf, Wp):
    """Returns (e', h_f, bits_f, h_p, bits_p); see _MPEdgeBlock."""
    return _MPEdgeBlock.apply(g, e, att, p_i, p_j, p_f, p_p, W0, W1, b1, W2, b2, Wf, Wp)


class _AttEdgeEncoder(torch.autograd.Function):
    """att_edge_encoder (clr_att_gnn.py:82-91, :164) 640 -> 512 -> 384 -> 256 -> 128 -> 64 in the pre-projected form
    (the two 288-wide node-side blocks of layer 0 arrive as row-gathered addends p_i[dst] + p_j[src]) as TWO fused
    launches: head 64 -> 512 -> 384 (the 512-wide tile alone is 128 KB of shared memory) and tail 384 -> 256 -> 128 ->
    64. Backward: one fused input-gradient chain 64 -> 128 -> 256 -> 384 (sign-bit masks), the 384 -> 512 step, then
    the weight gradients and the two per-node addend sums."""

    @staticmethod
    def forward(ctx, g, e0, p_i, p_j, W0, W1, b1, W2, b2, W3, b3, W4, b4):
        M, dev, bf = e0.size(0), e0.device, torch.bfloat16
        train = any(ctx.needs_input_grad)
        mk = lambda n: torch.empty((M, n), dtype=bf, device=dev)
        bt = lambda n: new_relu_bits(M, n, dev) if train else None
        x0 = mk(W0.size(0)) if train else None
        x1, x2, x3 = mk(W1.size(0)), (mk(W2.size(0)) if train else None), (mk(W3.size(0)) if train else None)
        y = mk(W4.size(0))
        bits = [bt(W0.size(0)), bt(W1.size(0)), bt(W2.size(0)), bt(W3.size(0))]
        head = [dict(W=W0, src=-1, act=L.ACT_RELU, adds=[(p_i, 0), (p_j, 1)], out=x0, bits_out=bits[0]),
                dict(W=W1, src=0, act=L.ACT_RELU, bias=b1, out=x1, bits_out=bits[1])]
        tail = [dict(W=W2, src=-1, act=L.ACT_RELU, bias=b2, out=x2, bits_out=bits[2]),
                dict(W=W3, src=0, act=L.ACT_RELU, bias=b3, out=x3, bits_out=bits[3]),
                dict(W=W4, src=1, act=L.ACT_NONE, bias=b4, out=y)]
        if not (chain_run([e0], head, g.by_dst.idx, g.by_src.idx, M) and chain_run([x1], tail, None, None, M)):
            raise RuntimeError("att_edge_encoder: chain plan rejected")
        ctx.g = g
        if train:
            ctx.save_for_backward(e0, x0, x1, x2, x3, *bits, W0, W1, W2, W3, W4)
        return y

    @staticmethod
    def backward(ctx, dy):
        e0, x0, x1, x2, x3, b0, b1_, b2_, b3_, W0, W1, W2, W3, W4 = ctx.saved_tensors
        g = ctx.g
        M, dev, bf = e0.size(0), e0.device, torch.bfloat16
        dy = dy if (dy.dtype == bf and _al16(dy)) else dy.to(bf).contiguous()
        mk = lambda n: torch.empty((M, n), dtype=bf, device=dev)
        dz3, dz2, dz1, dz0 = mk(W3.size(0)), mk(W2.size(0)), mk(W1.size(0)), mk(W0.size(0))
        tail = [dict(W=W4, transpose=True, src=-1, act=L.ACT_MASKBITS, bits_in=b3_, out=dz3),
                dict(W=W3, transpose=True, src=0, act=L.ACT_MASKBITS, bits_in=b2_, out=dz2),
                dict(W=W2, transpose=True, src=1, act=L.ACT_MASKBITS, bits_in=b1_, out=dz1)]
        head = [dict(W=W1, transpose=True, src=-1, act=L.ACT_MASKBITS, bits_in=b0, out=dz0)]
        if not (chain_run([dy], tail, None, None, M) and chain_run([dz1], head, None, None, M)):
            raise RuntimeError("att_edge_encoder backward: chain plan rejected")
        it = lambda t: (t, None, None, 0)
        d_e0 = linear_raw([it(dz0)], W0, None, M, trans_w=True, tc=True, out_dtype=bf) if ctx.needs_input_grad[1] else None
        dW0, _ = wgrad_raw(it(dz0), [it(e0)], M, W0.size(0), W0.size(1), want_bias=False, tc=True)
        dW1, db1 = wgrad_raw(it(dz1), [it(x0)], M, W1.size(0), W1.size(1), tc=True)
        dW2, db2 = wgrad_raw(it(dz2), [it(x1)], M, W2.size(0), W2.size(1), tc=True)
        dW3, db3 = wgrad_raw(it(dz3), [it(x2)], M, W3.size(0), W3.size(1), tc=True)
        dW4, db4 = wgrad_raw(it(dy), [it(x3)], M, W4.size(0), W4.size(1), tc=True)
        dp_i = segment_sum_raw(dz0, g.by_dst, out_dtype=bf)
        dp_j = segment_sum_raw(dz0, g.by_src, out_dtype=bf)
        return (None, d_e0, dp_i, dp_j, dW0, dW1, db1, dW2, db2, dW3, db3, dW4, db4)


def att_edge_encoder_supported(e0, linears):
    dims = [m.out_features for m in linears]
    return (_USE_CHAIN and _PRECISION == "bf16" and e0.size(0) >= _TC_MIN_ROWS and e0.dtype == torch.bfloat16 and _al16(e0)
            and e0.size(1) % 64 == 0 and len(linears) == 5 and dims == [512, 384, 256, 128, 64])


def att_edge_encoder_block(g, e0, p_i, p_j, w0_edge, linears):
    """linears: the five nn.Linear of att_edge_encoder; w0_edge = the edge-feature column block of layer 0."""
    l = linears
    return _AttEdgeEncoder.apply(g, e0, p_i, p_j, w0_edge, l[1].weight, l[1].bias, l[2].weight, l[2].bias,
                                 l[3].weight, l[3].bias, l[4].weight, l[4].bias)


class _NarrowMLP(torch.autograd.Function):
    """nn.Sequential(Linear, ReLU, ..., Linear[, Sigmoid]) with every width <= 64 as ONE kernel per
    direction (narrow_mlp.cu): hidden activations never leave registers; backward recomputes them."""

    @staticmethod
    def forward(ctx, nl, final_act, out_dtype, in_relu, x, *params):
        Ws = [w.contiguous() for w in params[0::2]]
        bs = [b.contiguous() if b is not None else None for b in params[1::2]]
        if x.dtype != torch.float64:           # float64 rows (edge_attr as stored) are cast inside the kernel
            x = _rows(x)
        dims = [Ws[0].size(1)] + [w.size(0) for w in Ws]
        M = x.size(0)
        y = torch.empty((M, dims[-1]), dtype=out_dtype, device=x.device)
        ctx.nl, ctx.dims, ctx.has_bias = nl, dims, [b is not None for b in bs]
        ctx.act = _ACT[final_act] | (L.NARROW_INPUT_RELU if in_relu else 0)
        if M > 0:
            L.check(L.lib().b3d_narrow_mlp_fwd(L.ptr(x), _DT[x.dtype], x.stride(0), M, nl, L.int_array(dims),
                                               L.ptr_array(Ws), L.ptr_array(bs), ctx.act, L.ptr(y), _DT[y.dtype],
                                               y.stride(0), L.stream()), "b3d_narrow_mlp_fwd")
        ctx.save_for_backward(x, *Ws, *[b for b in bs if b is not None])
        return y

    @staticmethod
    def backward(ctx, dy):
        nl, dims = ctx.nl, ctx.dims
        x, Ws = ctx.saved_tensors[0], list(ctx.saved_tensors[1:1 + nl])
        rest = list(ctx.saved_tensors[1 + nl:])
        bs = [rest.pop(0) if hb else None for hb in ctx.has_bias]
        dy = _rows(dy)
        if not dy.is_contiguous():
            dy = dy.contiguous()
        M = x.size(0)
        dX = torch.empty_like(x) if ctx.needs_input_grad[4] else None
        dWs = [torch.empty_like(w) for w in Ws]
        dbs = [torch.empty_like(b) if b is not None else None for b in bs]
        if M == 0:
            for t in dWs + [d for d in dbs if d is not None]:
                t.zero_()
            if dX is not None:
                dX.zero_()
        else:
            lib = L.lib()
            d32 = L.int_array(dims)
            wsb = int(lib.b3d_narrow_mlp_bwd_workspace_bytes(M, nl, d32))
            ws = torch.empty(wsb, dtype=torch.uint8, device=x.device)
            L.check(lib.b3d_narrow_mlp_bwd(L.ptr(x), _DT[x.dtype], x.stride(0), M, nl, d32, L.ptr_array(Ws),
                                           L.ptr_array(bs), ctx.act, L.ptr(dy), _DT[dy.dtype], dy.stride(0),
                                           L.ptr(dX), _DT[dX.dtype] if dX is not None else 0,
                                           dX.stride(0) if dX is not None else 0, L.ptr_array(dWs), L.ptr_array(dbs),
                                           0, L.ptr(ws), wsb, L.stream()), "b3d_narrow_mlp_bwd")
        flat = []
        for dw, db in zip(dWs, dbs):
            flat += [dw, db]
        return (None, None, None, None, dX, *flat)


def _narrow_dims_ok(x, dims):
    if x.dim() == 2 and x.dtype == torch.float64 and x.stride(1) == 1 and not x.requires_grad and \
            x.data_ptr() % 16 == 0 and x.stride(0) % 2 == 0:
        pass                                   # float64 input rows: loaded as double2, cast in registers
    elif x.dim() != 2 or x.dtype not in (torch.float32, torch.bfloat16) or x.stride(1) != 1 or not _al16(x):
        return False
    return x.size(1) == dims[0] and bool(L.lib().b3d_narrow_mlp_supported(len(dims) - 1, L.int_array(dims)))


def _narrow_split(inputs, weights, biases, final_act, row_mask, adds, out_dtype, premasked):
    """bf16 mode: the widest layer of a narrow chain as a tensor-core layer, the rest as the narrow kernel —
    classifier 64 -> 32 | ReLU, 32 -> 16 -> 8 -> 1 [Sigmoid] (clr_att_gnn.py:49-57) and edge encoder
    4 -> 16 -> 32, ReLU | 32 -> 64 (clr_att_gnn.py:35-41): three quarters of either chain's multiply-adds leave the
    FFMA pipe (narrow_mlp.cu ran at 24 TFLOP/s: 7 % of the step), at the price of one bf16 [rows, 32] round trip.
    Returns None when the chain is not one of the two."""
    if _PRECISION != "bf16" or not FEATURES["narrow_split"] or len(inputs) != 1 or inputs[0][1] is not None \
            or row_mask is not None or adds or premasked:
        return None
    x = inputs[0][0]
    dims = [weights[0].size(1)] + [w.size(0) for w in weights]
    if x.size(0) < _TC_MIN_ROWS:
        return None
    bf = torch.bfloat16
    if dims == [64, 32, 16, 8, 1] and final_act in (None, "sigmoid") and x.dtype == bf and _al16(x) and x.stride(1) == 1:
        z0 = _FusedMLP.apply(1, None, None, (None,), (), bf, False, weights[0], biases[0], x)
        if not _narrow_dims_ok(z0, dims[1:]):
            return None
        flat = []
        for w, b in zip(weights[1:], biases[1:]):
            flat += [w, b]
        return _NarrowMLP.apply(3, final_act, out_dtype or torch.float32, True, z0, *flat)
    if dims == [4, 16, 32, 64] and final_act is None and _narrow_dims_ok(x, dims[:3]):
        flat = []
        for w, b in zip(weights[:2], biases[:2]):
            flat += [w, b]
        h = _NarrowMLP.apply(2, "relu", bf, False, x, *flat)
        return _FusedMLP.apply(1, None, None, (None,), (), out_dtype, False, weights[2], biases[2], h)
    return None


def _narrow_ok(inputs, weights, final_act, row_mask, adds, premasked):
    if len(inputs) != 1 or inputs[0][1] is not None or row_mask is not None or adds or premasked:
        return False
    if final_act not in (None, "sigmoid") or len(weights) not in (3, 4) or (final_act and len(weights) != 4):
        return False
    dims = [weights[0].size(1)] + [w.size(0) for w in weights]
    return _narrow_dims_ok(inputs[0][0], dims)


def fused_mlp(inputs, weights, biases, final_act=None, row_mask=None, adds=(), out_dtype=None, premasked=False):
    """inputs: list of (tensor [rows,w], NodeIndex|None); the concatenation order defines the first
    weight's input-column layout. weights/biases: per layer.
    adds: up to two (tensor [n, out_features of layer 0] fp32, NodeIndex|None) summed, row-gathered,
    into layer 0's pre-activation (node-side weight blocks applied per node)."""
    xs = [t for t, _ in inputs]
    nidx = tuple(ni for _, ni in inputs)
    flat = []
    for w, b in zip(weights, biases):
        flat += [w, b]
    y = _narrow_split(inputs, weights, biases, final_act, row_mask, adds, out_dtype, premasked)
    if y is not None:
        return y
    if _narrow_ok(inputs, weights, final_act, row_mask, adds, premasked):
        return _NarrowMLP.apply(len(weights), final_act, out_dtype or torch.float32, False, xs[0], *flat)
    return _FusedMLP.apply(len(weights), final_act, row_mask, nidx, tuple(ni for _, ni in adds), out_dtype,
                           premasked, *flat, *xs, *[t for t, _ in adds])


def edge_attr_rows(edge_attr):
    """edge_attr as the edge encoder's input: the stored float64 [E,4] matrix goes to the narrow-chain kernel
    as is (its `.float()` cast, pose_gnn.py:67 / clr_att_gnn.py:123, happens in the row load); anything
    else is cast here."""
    if edge_attr.dtype == torch.float64 and edge_attr.dim() == 2 and edge_attr.is_contiguous() and \
            not edge_attr.requires_grad:
        return edge_attr
    return edge_attr.float()


def fused_linear(inputs, weight, bias=None, act=None, row_mask=None, adds=()):
    """Single layer: act(cat(inputs) W^T + b [+ gathered addends])."""
    return fused_mlp(inputs, [weight], [bias], final_act=act, row_mask=row_mask, adds=adds)


def run_mlp(seq, inputs, final_act=None, row_mask=None, out_dtype=None):
    """Run an nn.Sequential of Linear/ReLU(/Sigmoid) parameter containers as one fused chain."""
    linears = [m for m in seq if isinstance(m, torch.nn.Linear)]
    return fused_mlp(inputs, [m.weight for m in linears], [m.bias for m in linears], final_act, row_mask,
                     out_dtype=out_dtype)


class _SegmentSum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, nidx, relu_src, relu_bits, out_dtype):
        ctx.nidx, ctx.src_dtype = nidx, src.dtype
        use_bits = relu_src and relu_bits is not None and relu_bits.numel() > 0 and src.dtype == torch.bfloat16
        ctx.use_bits = use_bits
        ctx.save_for_backward(relu_bits if use_bits else (src if relu_src else None))
        return segment_sum_raw(src, nidx, out_dtype=out_dtype or torch.float32)

    @staticmethod
    def backward(ctx, dout):
        (mask,) = ctx.saved_tensors
        if ctx.use_bits:
            g = gather_rows_raw(dout, ctx.nidx.idx, out_dtype=ctx.src_dtype, relu_bits=mask)
        else:
            g = gather_rows_raw(dout, ctx.nidx.idx, out_dtype=ctx.src_dtype, relu_mask=mask)
        return (g if g.dtype == ctx.src_dtype else g.to(ctx.src_dtype)), None, None, None, None


def segment_sum(src, nidx, relu_src=False, relu_bits=None, out_dtype=None):
    """out[n] = sum of src rows whose endpoint is n (torch_scatter.scatter(reduce='add')).
    relu_src=True: `src` is the output of a ReLU layer built with premasked=True; the backward then
    returns the gradient already multiplied by (src > 0), fused into the row gather (from the
    layer's sign bits `relu_bits` when given, else from `src` itself)."""
    return _SegmentSum.apply(src, nidx, relu_src, relu_bits, out_dtype)


def add_n_raw(ts, out_dtype=None):
    """sum of equally shaped [M,C] tensors (fp32/bf16, row strides free) in ONE kernel."""
    ts = [_rows(t) for t in ts]
    M, C_ = ts[0].shape
    out = torch.empty((M, C_), dtype=out_dtype or ts[0].dtype, device=ts[0].device)
    if M == 0:
        return out
    ok = C_ % 8 == 0 and all(_al16(t) for t in ts) and len(ts) <= L.MAX_SEGS
    if not ok:                      # shapes the kernel does not take (never on the model path)
        acc = ts[0].float()
        for t in ts[1:]:
            acc = acc + t.float()
        return acc.to(out.dtype)
    L.check(L.lib().b3d_add_n(L.make_segs([(t, None, None, 0) for t in ts]), len(ts), M, L.ptr(out), _DT[out.dtype],
                              out.stride(0), L.stream()), "b3d_add_n")
    return out


class _Fanout(torch.autograd.Function):
    """Identity with n outputs. A tensor with several consumers gets its gradient as ONE n-way sum
    kernel (b3d_add_n) instead of n-1 pairwise additions by the autograd engine (which also fall on
    the slow strided path when a gradient is a column slice of a wider matrix)."""

    @staticmethod
    def forward(ctx, x, n):
        ctx.dtype = x.dtype
        return tuple(x.view_as(x) for _ in range(n))

    @staticmethod
    def backward(ctx, *gs):
        gs = [g for g in gs if g is not None]
        if not gs:
            return None, None
        if len(gs) == 1:
            g = gs[0]
            return (g if g.dtype == ctx.dtype else g.to(ctx.dtype)), None
        return add_n_raw(gs, out_dtype=ctx.dtype), None


def fanout(x, n):
    """n aliases of x whose gradients are summed by one kernel."""
    if not (torch.is_grad_enabled() and x.requires_grad) or n < 2:
        return (x,) * n
    return _Fanout.apply(x, n)


class _SplitCols(torch.autograd.Function):
    """Column split x -> (x[:, :s0], x[:, s0:s0+s1], ...) whose backward is ONE concatenation instead
    of a zero-filled full-width gradient plus an addition per slice."""

    @staticmethod
    def forward(ctx, x, sizes):
        ctx.sizes, ctx.dtype, ctx.rows = sizes, x.dtype, x.size(0)
        outs, off = [], 0
        for w in sizes:
            outs.append(x[:, off:off + w])
            off += w
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gs):
        ref = next(g for g in gs if g is not None)
        parts = [g.to(ctx.dtype) if g is not None else ref.new_zeros((ctx.rows, w), dtype=ctx.dtype)
                 for g, w in zip(gs, ctx.sizes)]
        return torch.cat(parts, 1), None


def split_cols(x, sizes):
    assert sum(sizes) == x.size(1)
    if not (torch.is_grad_enabled() and x.requires_grad):
        outs, off = [], 0
        for w in sizes:
            outs.append(x[:, off:off + w])
            off += w
        return tuple(outs)
    return _SplitCols.apply(x, tuple(sizes))


class _BCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inp, y, w, scale, from_logits, focal=None):
        x = inp.reshape(-1).contiguous().float()
        E = x.numel()
        if E == 0:      # a batch whose windows hold no edges: mean over nothing, like torch's BCELoss (nan, no gradient)
            ctx.save_for_backward(torch.empty_like(x))
            ctx.shape = inp.shape
            return x.new_full((), float("nan"))
        lib = L.lib()
        loss = torch.empty(1, dtype=torch.float32, device=x.device)
        grad = torch.empty_like(x)
        part = torch.empty(int(lib.b3d_bce_partials(E)), dtype=torch.float32, device=x.device)
        yy = y.reshape(-1).to(torch.int64).contiguous()
        ww = w.reshape(-1).contiguous().float() if w is not None else None
        if focal is not None:
            L.check(lib.b3d_focal_fwd_bwd(L.ptr(x), L.ptr(yy), L.ptr(ww), E, float(scale), int(from_logits),
                                          float(focal[0]), float(focal[1]), L.ptr(loss), L.ptr(grad), L.ptr(part),
                                          L.stream()), "b3d_focal_fwd_bwd")
        else:
            L.check(lib.b3d_bce_fwd_bwd(L.ptr(x), L.ptr(yy), L.ptr(ww), E, float(scale), int(from_logits),
                                        L.ptr(loss), L.ptr(grad), L.ptr(part), L.stream()), "b3d_bce_fwd_bwd")
        ctx.save_for_backward(grad)
        ctx.shape = inp.shape
        return loss.reshape(())

    @staticmethod
    def backward(ctx, dl):
        (grad,) = ctx.saved_tensors
        return (grad * dl).reshape(ctx.shape), None, None, None, None, None


def bce_loss(out, y, weight=None, batch_size=1, from_logits=False):
    """BCELoss(weight)(out, y) / batch_size (train.py:136-141); from_logits for PoseGNN (C11)."""
    return _BCE.apply(out, y, weight, 1.0 / batch_size, from_logits)


def focal_loss(out, y, weight=None, batch_size=1, alpha=0.25, gamma=2.0, from_logits=False):
    """Focal edge loss (BASELINE config 5 "focal/BCE"; not part of the reference, see b3d.h):
    mean_e w_e * alpha_t (1 - p_t)^gamma (-log p_t) / batch_size."""
    return _BCE.apply(out, y, weight, 1.0 / batch_size, from_logits, (alpha, gamma))
