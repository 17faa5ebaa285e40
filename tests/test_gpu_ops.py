"""GPU parity tests of the individual libb3d entry points (called through the C ABI via
ctypes) against the CPU oracle. Integer / index work is bit-exact; fp32 work uses the
tolerance stated in each test (north_star: 1e-4 relative for the fp32 path)."""
from types import SimpleNamespace

import pytest
import torch

from oracle import ref_restated as R
from batch3dmot_b200 import _lib as L, ops, synth
from .conftest import load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(autouse=True)
def exact_mode():
    """Kernel-level tests of the plain-fp32 (FFMA) entry points against float64: the "exact" arithmetic mode.
    (The default "fp32" mode runs the same layers as split-bf16 tensor-core tiles: tests/test_gpu_split.py.)"""
    ops.set_precision("exact")
    yield
    ops.set_precision("fp32")


def rel_err(a, b):
    return float((a.detach().double().cpu() - b.detach().double().cpu()).abs().max() / b.double().abs().max().clamp_min(1e-30))


# ------------------------------------------------------------------ CSR / radix sort (bit-exact)
@pytest.mark.parametrize("E,N", [(0, 5), (1, 1), (1000, 50), (60489, 2000), (300000, 70000), (5000, 3)])
def test_csr_build_bit_exact(E, N):
    g = torch.Generator().manual_seed(E + N)
    ei = torch.randint(0, N, (2, E), generator=g)
    G = ops.Graph(ei.to(DEV), N, check=True)
    for idx, rowptr, perm in ((ei[1], G.rowptr_dst, G.perm_dst), (ei[0], G.rowptr_src, G.perm_src)):
        rp, pm = R.csr_build(idx, N)
        assert torch.equal(rowptr.cpu().long(), rp)
        assert torch.equal(perm.cpu().long(), pm)
    assert torch.equal(G.src32.cpu().long(), ei[0]) and torch.equal(G.dst32.cpu().long(), ei[1])


def test_csr_reference_ordering_and_bad_index():
    g = synth.scene_graph(seed=5621)
    G = ops.Graph(g.edge_index.to(DEV), g.num_nodes, check=True)
    # reference edge order is target-sorted -> the CSC permutation is the identity
    assert torch.equal(G.perm_dst.cpu().long(), torch.arange(G.E))
    bad = g.edge_index.clone(); bad[0, 7] = g.num_nodes
    with pytest.raises(IndexError):
        ops.Graph(bad.to(DEV), g.num_nodes, check=True)


def test_csr_large_full_size_properties():
    """BASELINE-sized batch (32 scenes): sortedness + permutation properties instead of the oracle."""
    scenes = [synth.scene_graph(seed=100 + i) for i in range(8)]
    c = synth.collate(scenes)
    ei = c.edge_index.to(DEV)
    G = ops.Graph(ei, c.num_nodes, check=True)
    for key, perm, rowptr in ((ei[0], G.perm_src, G.rowptr_src), (ei[1], G.perm_dst, G.rowptr_dst)):
        p = perm.long()
        assert torch.equal(torch.sort(p).values, torch.arange(G.E, device=DEV))       # a permutation
        ks = key[p]
        assert bool((ks[1:] >= ks[:-1]).all())                                         # sorted
        same = ks[1:] == ks[:-1]
        assert bool((p[1:][same] > p[:-1][same]).all())                                # stable
        assert torch.equal(rowptr.long()[1:] - rowptr.long()[:-1], torch.bincount(key, minlength=c.num_nodes))


# ------------------------------------------------------------------ segmented sum (bit-exact vs index_add_)
@pytest.mark.parametrize("C", [64, 128, 192, 48, 5])
def test_segment_sum_bit_exact(C):
    g = synth.scene_graph(seed=7, T=12, nodes_per_frame=30)
    ei, N = g.edge_index, g.num_nodes
    G = ops.Graph(ei.to(DEV), N)
    src = torch.randn(ei.size(1), C)
    for nidx, index in ((G.by_dst, ei[1]), (G.by_src, ei[0])):
        got = ops.segment_sum_raw(src.to(DEV), nidx)
        ref = R.scatter_add(src, index, N)          # sequential CPU order
        assert torch.equal(got.cpu(), ref)
    # strided source (column slice) + accumulate
    wide = torch.randn(ei.size(1), C + 8).to(DEV)
    out = torch.ones(N, C, device=DEV)
    ops.segment_sum_raw(wide[:, 4:4 + C], G.by_src, out=out, accumulate=True)
    ref = 1 + R.scatter_add(wide[:, 4:4 + C].cpu(), ei[0], N)
    assert rel_err(out, ref) < 1e-6


def test_segment_sum_empty_and_isolated():
    ei = torch.tensor([[0, 0, 3], [5, 5, 5]])
    G = ops.Graph(ei.to(DEV), 8)
    src = torch.randn(3, 64)
    got = ops.segment_sum_raw(src.to(DEV), G.by_dst).cpu()
    assert torch.equal(got[5], (src[0] + src[1]) + src[2]) and float(got[[0, 1, 2, 3, 4, 6, 7]].abs().max()) == 0
    assert torch.equal(ops.gather_rows_raw(got.to(DEV), G.dst32).cpu(), got[ei[1]])


@pytest.mark.parametrize("C", [7, 64, 192])
def test_gather_rows_fp32_with_relu_mask(C):
    """fp32 gather with the fp32 ReLU activation as mask (backward of an fp32 segment sum) in one pass."""
    torch.manual_seed(C)
    N, M = 333, 10007
    src, act = torch.randn(N, C), torch.randn(M, C)
    idx = torch.randint(0, N, (M,), dtype=torch.int32)
    before = L.launch_count()
    got = ops.gather_rows_raw(src.to(DEV), idx.to(DEV), relu_mask=act.to(DEV))
    assert L.launch_count() - before == 1
    assert torch.equal(got.cpu(), src[idx.long()] * (act > 0))


@pytest.mark.parametrize("C", [64, 96, 192, 256, 512])
@pytest.mark.parametrize("src_bf16", [True, False])
def test_gather_rows_bf16_with_relu_masks(C, src_bf16):
    """Backward of a bf16 segment sum: out[r] = src[idx[r]] (rounded to bf16), zeroed where the ReLU mask says so; the
    mask as sign bits (B3D_BITS words) or as the bf16 activation itself; M not a multiple of the rows per warp."""
    torch.manual_seed(C)
    N, M = 333, 10007
    src = torch.randn(N, C)
    src = src.to(torch.bfloat16) if src_bf16 else src
    idx = torch.randint(0, N, (M,), dtype=torch.int32)
    act = torch.randn(M, C).to(torch.bfloat16)
    keep = act > 0
    words = (keep.reshape(M, C // 32, 32).to(torch.int64) << torch.arange(32)).sum(2).t().contiguous()
    bits = torch.where(words >= 2 ** 31, words - 2 ** 32, words).to(torch.int32).to(DEV)
    ref = src[idx.long()].to(torch.bfloat16)
    got = ops.gather_rows_raw(src.to(DEV), idx.to(DEV), out_dtype=torch.bfloat16)
    assert torch.equal(got.cpu(), ref)
    got = ops.gather_rows_raw(src.to(DEV), idx.to(DEV), out_dtype=torch.bfloat16, relu_bits=bits)
    assert torch.equal(got.cpu(), ref * keep)
    got = ops.gather_rows_raw(src.to(DEV), idx.to(DEV), out_dtype=torch.bfloat16, relu_mask=act.to(DEV))
    assert torch.equal(got.cpu(), ref * keep)


# ------------------------------------------------------------------ dense layers
@pytest.mark.parametrize("widths,n_out,gather", [((48, 48, 32), 96, True), ((96, 96, 64, 64), 256, True),
                                                ((19,), 24, False), ((4,), 8, False), ((64,), 1, False),
                                                ((64, 128, 96, 64, 128, 96, 64), 512, True)])
def test_linear_forward(widths, n_out, gather):
    torch.manual_seed(sum(widths) + n_out)
    N, M = 300, 1111
    idx = torch.randint(0, N, (len(widths), M))
    xs = [torch.randn(N if gather and s % 3 != 2 else M, w) for s, w in enumerate(widths)]
    W, b = torch.randn(n_out, sum(widths)) * 0.1, torch.randn(n_out)
    cat = torch.cat([x[idx[s]] if x.size(0) == N and gather else x for s, x in enumerate(xs)], 1)
    items = [(x.to(DEV), idx[s].int().to(DEV) if (x.size(0) == N and gather) else None, None, 0)
             for s, x in enumerate(xs)]
    for act, fn in ((L.ACT_NONE, lambda t: t), (L.ACT_RELU, torch.relu), (L.ACT_SIGMOID, torch.sigmoid)):
        y = ops.linear_raw(items, W.to(DEV), b.to(DEV), M, act)
        ref = fn(cat.double() @ W.double().t() + b.double())
        assert rel_err(y, ref) < 2e-6          # fp32 FMA accumulation vs float64 reference
    # transposed weight (input gradient) + output ReLU mask + accumulate
    dy = torch.randn(M, n_out)
    hmask = torch.randn(M, sum(widths))
    out = torch.full((M, sum(widths)), 0.5, device=DEV)
    ops.linear_raw([(dy.to(DEV), None, None, 0)], W.to(DEV), None, M, trans_w=True, out=out, accumulate=True,
                   out_mask=hmask.to(DEV))
    ref = 0.5 + (dy.double() @ W.double()) * (hmask > 0)
    assert rel_err(out, ref) < 2e-6


def test_linear_row_mask_and_operand_masks():
    torch.manual_seed(3)
    M, K, n_out = 500, 40, 24
    x, W = torch.randn(M, K), torch.randn(n_out, K)
    rm = (torch.rand(M) < 0.7)
    y = ops.linear_raw([(x.to(DEV), None, None, 0)], W.to(DEV), None, M, row_mask=rm.to(torch.uint8).to(DEV))
    assert rel_err(y, (x @ W.t()) * rm[:, None]) < 2e-6
    act = torch.rand(M, K) - 0.3
    y = ops.linear_raw([(x.to(DEV), None, act.to(DEV), L.MASK_RELU)], W.to(DEV), None, M)
    assert rel_err(y, ((x * (act > 0)).double() @ W.double().t())) < 2e-6
    sg = torch.sigmoid(act)
    y = ops.linear_raw([(x.to(DEV), None, sg.to(DEV), L.MASK_SIGMOID)], W.to(DEV), None, M)
    assert rel_err(y, ((x * sg * (1 - sg)).double() @ W.double().t())) < 2e-6


@pytest.mark.parametrize("widths,n_out,M", [((48, 48, 32), 96, 3000), ((19,), 24, 700), ((96, 64, 96), 192, 20000),
                                            ((4,), 1, 10), ((96, 96, 64, 64), 256, 70000)])
def test_wgrad(widths, n_out, M):
    torch.manual_seed(M)
    N = 257
    idx = torch.randint(0, N, (len(widths), M))
    xs = [torch.randn(N if s % 2 == 0 else M, w) for s, w in enumerate(widths)]
    cat = torch.cat([x[idx[s]] if x.size(0) == N else x for s, x in enumerate(xs)], 1).double()
    dy, y = torch.randn(M, n_out), torch.randn(M, n_out)
    items = [(x.to(DEV), idx[s].int().to(DEV) if x.size(0) == N else None, None, 0) for s, x in enumerate(xs)]
    dW, db = ops.wgrad_raw((dy.to(DEV), None, y.to(DEV), L.MASK_RELU), items, M, n_out, sum(widths))
    dym = (dy * (y > 0)).double()
    assert rel_err(dW, dym.t() @ cat) < 1e-5 and rel_err(db, dym.sum(0)) < 1e-5
    # accumulate + determinism
    dW2, db2 = ops.wgrad_raw((dy.to(DEV), None, y.to(DEV), L.MASK_RELU), items, M, n_out, sum(widths),
                             dW=dW.clone(), db=db.clone(), accumulate=True)
    assert rel_err(dW2, 2 * (dym.t() @ cat)) < 1e-5
    dW3, _ = ops.wgrad_raw((dy.to(DEV), None, y.to(DEV), L.MASK_RELU), items, M, n_out, sum(widths))
    assert torch.equal(dW3, dW)


def test_fused_linear_autograd_matches_torch():
    torch.manual_seed(5)
    g = synth.scene_graph(seed=9, T=8, nodes_per_frame=15, k=8)
    G = ops.Graph(g.edge_index.to(DEV), g.num_nodes)
    N, E = g.num_nodes, g.edge_index.size(1)
    x = torch.randn(N, 48, requires_grad=True); e = torch.randn(E, 32, requires_grad=True)
    W = (torch.randn(96, 128) * 0.1).requires_grad_(True); b = torch.randn(96, requires_grad=True)
    ref = torch.relu(torch.cat([x[g.edge_index[1]], x[g.edge_index[0]], e], 1) @ W.t() + b)
    m_ref = R.scatter_add(ref, g.edge_index[0], N)
    (m_ref ** 2).sum().backward()
    xc, ec, Wc, bc = [t.detach().to(DEV).requires_grad_(True) for t in (x, e, W, b)]
    y = ops.fused_linear([(xc, G.by_dst), (xc, G.by_src), (ec, None)], Wc, bc, "relu")
    m = ops.segment_sum(y, G.by_src)
    (m ** 2).sum().backward()
    assert rel_err(m, m_ref) < 1e-5
    for a, r in ((xc, x), (ec, e), (Wc, W), (bc, b)):
        assert rel_err(a.grad, r.grad) < 1e-4


# ------------------------------------------------------------------ k-NN (bit-exact) + GAT
def test_knn_golden_and_gat():
    g = load_golden("knn_gat_small.pt")
    idx = ops.knn_frames(g["x"].to(DEV), g["frame_ptr"], g["k"])
    assert torch.equal(idx.cpu(), g["idx"])
    sd = g["gat_state_dict"]
    h = ops.linear_raw([(g["x"].to(DEV), None, None, 0)], sd["knn_conv.lin_src.weight"].to(DEV), None, g["x"].size(0))
    out = ops.gat_aggregate(h, sd["knn_conv.att_src"].to(DEV), sd["knn_conv.att_dst"].to(DEV),
                            sd["knn_conv.bias"].to(DEV), idx)
    assert rel_err(out, g["gat_out"]) < 1e-5


@pytest.mark.parametrize("N,D,k,frame", [(600, 48, 8, 50), (900, 96, 20, 30), (40, 48, 20, 7)])
def test_gat_backward_matches_oracle_autograd(N, D, k, frame):
    """b3d_gat_bwd (softmax / LeakyReLU / score paths + deterministic reversed-table reduction) against
    torch autograd through the oracle's GATConv restatement; frames smaller than k+1 exercise the -1 padding."""
    torch.manual_seed(N + k)
    x = torch.randn(N, D)
    ptr = torch.arange(0, N + 1, frame)
    if ptr[-1] != N:
        ptr = torch.cat([ptr, torch.tensor([N])])
    idx = R.knn_frames(x, ptr, k)
    params = {"knn_conv.lin_src.weight": (torch.randn(D, D) * 0.2).requires_grad_(True),
              "knn_conv.att_src": (torch.randn(1, 1, D) * 0.3).requires_grad_(True),
              "knn_conv.att_dst": (torch.randn(1, 1, D) * 0.3).requires_grad_(True),
              "knn_conv.bias": torch.randn(D).requires_grad_(True)}
    xr = x.clone().requires_grad_(True)
    out_ref = R.gat_conv(params, xr, R.knn_to_edge_index(idx))
    gout = torch.randn(N, D)
    out_ref.backward(gout)
    from batch3dmot_b200.gat import GATConv
    conv = GATConv(D, D).to(DEV)
    with torch.no_grad():
        conv.lin_src.weight.copy_(params["knn_conv.lin_src.weight"]); conv.att_src.copy_(params["knn_conv.att_src"])
        conv.att_dst.copy_(params["knn_conv.att_dst"]); conv.bias.copy_(params["knn_conv.bias"])
    xd = x.to(DEV).requires_grad_(True)
    out = conv.forward_table(xd, idx.to(DEV))
    assert rel_err(out, out_ref) < 1e-5
    out.backward(gout.to(DEV))
    assert rel_err(xd.grad, xr.grad) < 1e-4
    assert rel_err(conv.lin_src.weight.grad, params["knn_conv.lin_src.weight"].grad) < 1e-4
    assert rel_err(conv.att_src.grad, params["knn_conv.att_src"].grad) < 1e-4
    assert rel_err(conv.att_dst.grad, params["knn_conv.att_dst"].grad) < 1e-4
    assert rel_err(conv.bias.grad, params["knn_conv.bias"].grad) < 1e-4
    out2 = conv.forward_table(xd, idx.to(DEV)); xd.grad = None; out2.backward(gout.to(DEV))
    g1 = xd.grad.clone(); out3 = conv.forward_table(xd, idx.to(DEV)); xd.grad = None; out3.backward(gout.to(DEV))
    assert torch.equal(g1, xd.grad), "the backward must be deterministic"


@pytest.mark.parametrize("N,frame,D,k,skewed", [(2000, 250, 48, 8, False), (3000, 250, 96, 16, False),
                                                (2500, 100, 48, 20, True), (600, 7, 96, 20, False),
                                                (1500, 300, 19, 32, False)])
def test_knn_grid_features_bit_exact(N, frame, D, k, skewed):
    """B.3: grid features make every distance exact in fp32; ties are frequent -> tests the
    (distance, index) tie-break, frames smaller than k, frames of one node."""
    x, ptr = synth.knn_stress(N + k, N, frame=frame, D=D, skewed=skewed)
    idx = ops.knn_frames(x.to(DEV), ptr, k)
    assert torch.equal(idx.cpu(), R.knn_frames(x, ptr, k))


def test_knn_continuous_features():
    x, ptr = synth.knn_stress(77, 2000, frame=200, D=48, continuous=True)
    idx = ops.knn_frames(x.to(DEV), ptr, 20).cpu()
    ref = R.knn_frames(x, ptr, 20)
    assert torch.equal(idx, ref)   # same op order (sub, mul, add in ascending d, no FMA) -> identical


def test_knn_large_properties():
    """Config-3 size (200k nodes): compare a sampled subset of queries with the oracle."""
    x, ptr = synth.knn_stress(5, 200000, frame=250, D=48)
    idx = ops.knn_frames(x.to(DEV), ptr, 16).cpu()
    assert int(idx.min()) >= 0
    fr = torch.arange(200000) // 250
    assert bool((fr[idx] == fr[:, None]).all())                       # frame-local
    assert bool((idx != torch.arange(200000)[:, None]).all())        # self excluded
    sub = slice(250 * 11, 250 * 13)
    assert torch.equal(idx[sub], R.knn_frames(x[sub], torch.tensor([0, 250, 500]), 16) + 250 * 11)


# ------------------------------------------------------------------ masks, loss, optimiser
def test_row_nonzero():
    f = torch.zeros(100, 128, 3)
    present = torch.rand(100) < 0.6
    f[present] = torch.randn(int(present.sum()), 128, 3)
    f[3] = 0; f[3, 5, 1] = 1e-30; present[3] = True
    assert torch.equal(ops.row_nonzero(f.to(DEV)).cpu(), present)


@pytest.mark.parametrize("from_logits", [False, True])
def test_bce(from_logits):
    torch.manual_seed(0)
    E = 10007
    z = torch.randn(E) * 3
    y = (torch.rand(E) < 0.1).long()
    w = torch.rand(E) + 0.5
    inp = (z if from_logits else torch.sigmoid(z)).requires_grad_(True)
    ref = (R.bce_logits_loss if from_logits else R.bce_loss)(inp, y, w, batch_size=2)
    ref.backward()
    ic = inp.detach().to(DEV).requires_grad_(True)
    loss = ops.bce_loss(ic, y.to(DEV), w.to(DEV), batch_size=2, from_logits=from_logits)
    loss.backward()
    assert abs(loss.item() - ref.item()) < 1e-5 * abs(ref.item())
    assert rel_err(ic.grad, inp.grad) < 1e-5


@pytest.mark.parametrize("from_logits,gamma,alpha", [(False, 2.0, 0.25), (True, 2.0, 0.25), (False, 0.0, 0.5), (True, 1.5, 0.75)])
def test_focal_loss(from_logits, gamma, alpha):
    """Focal loss (not in the reference: published definition restated in oracle.ref_restated.focal_loss);
    gamma = 0, alpha = 0.5 must reduce to half the BCE."""
    torch.manual_seed(4)
    E = 10007
    z = torch.randn(E).double() * 3
    y = (torch.rand(E) < 0.1).long()
    w = (torch.rand(E) + 0.5).double()
    # the reference sees the SAME fp32-rounded inputs (1 - p loses digits for p near 1 otherwise)
    inp = (z if from_logits else torch.sigmoid(z)).float().double().requires_grad_(True)
    ref = R.focal_loss(inp, y, w, batch_size=2, alpha=alpha, gamma=gamma, from_logits=from_logits)
    ref.backward()
    ic = inp.detach().float().to(DEV).requires_grad_(True)
    loss = ops.focal_loss(ic, y.to(DEV), w.float().to(DEV), batch_size=2, alpha=alpha, gamma=gamma, from_logits=from_logits)
    loss.backward()
    assert abs(loss.item() - ref.item()) < 2e-5 * abs(ref.item())
    assert rel_err(ic.grad, inp.grad) < 2e-5
    if gamma == 0.0:
        bce = R.bce_loss(inp.detach().float(), y, w.float(), batch_size=2)
        assert abs(loss.item() - 0.5 * bce.item()) < 2e-5 * abs(bce.item())


def test_adam_matches_torch():
    torch.manual_seed(1)
    p = torch.randn(5000, requires_grad=True)
    opt = torch.optim.Adam([p], lr=1e-4, weight_decay=1e-4, betas=(0.9, 0.999))
    pc = p.detach().clone().to(DEV); m = torch.zeros_like(pc); v = torch.zeros_like(pc)
    for step in range(1, 4):
        g = torch.randn(5000)
        p.grad = g.clone()
        opt.step()
        ops.adam_step(pc, g.to(DEV), m, v, 1e-4, (0.9, 0.999), 1e-8, 1e-4, step)
    assert float((pc.cpu() - p.detach()).abs().max()) < 5e-7   # a few ulp: fused vs foreach op order


def _torch_chain(dims, final_sigmoid):
    mods = []
    for i in range(len(dims) - 1):
        mods.append(torch.nn.Linear(dims[i], dims[i + 1]))
        if i + 2 < len(dims):
            mods.append(torch.nn.ReLU())
    if final_sigmoid:
        mods.append(torch.nn.Sigmoid())
    return torch.nn.Sequential(*mods)


@pytest.mark.parametrize("dims,sig", [((4, 16, 32, 64), False), ((4, 8, 16, 32), False), ((64, 32, 16, 8, 1), True),
                                      ((64, 32, 16, 8, 1), False), ((32, 16, 8, 4, 1), False)])
@pytest.mark.parametrize("M", [1, 255, 1000, 70001])
def test_narrow_mlp_chain_matches_torch(dims, sig, M):
    """narrow_mlp.cu (whole edge-encoder / edge-classifier chain in one kernel per direction) against
    torch fp32 autograd: forward 1e-5, input and parameter gradients 1e-4 of each tensor's max."""
    torch.manual_seed(M + dims[1])
    seq = _torch_chain(dims, sig).double()
    x = torch.randn(M, dims[0]).double().requires_grad_(True)
    y_ref = seq(x)
    gy = torch.randn_like(y_ref)
    y_ref.backward(gy)
    dev_seq = _torch_chain(dims, sig).to(DEV)
    dev_seq.load_state_dict({k: v.float() for k, v in seq.state_dict().items()})
    xd = x.detach().float().to(DEV).requires_grad_(True)
    before = L.launch_count()
    y = ops.run_mlp(dev_seq, [(xd, None)], final_act="sigmoid" if sig else None)
    assert L.launch_count() - before == 1, "the chain must be ONE k_narrow_fwd launch"
    assert y.dtype == torch.float32 and rel_err(y, y_ref) < 1e-5
    before = L.launch_count()
    y.backward(gy.float().to(DEV))
    assert L.launch_count() - before == 2, "k_narrow_bwd + k_narrow_reduce"
    assert rel_err(xd.grad, x.grad) < 1e-4
    for (n, p), (_, q) in zip(dev_seq.named_parameters(), seq.named_parameters()):
        assert rel_err(p.grad, q.grad) < 1e-4, n


def test_narrow_mlp_bf16_storage():
    """bf16 X / Y / dY / dX storage around the fp32 chain (the bf16 mode's edge tensors)."""
    torch.manual_seed(3)
    M = 5000
    dims = (64, 32, 16, 8, 1)
    seq = _torch_chain(dims, True).to(DEV)
    x16 = torch.randn(M, 64).to(torch.bfloat16).to(DEV).requires_grad_(True)
    y = ops.run_mlp(seq, [(x16, None)], final_act="sigmoid")
    x32 = x16.detach().float().requires_grad_(True)
    y_ref = seq(x32)
    assert rel_err(y, y_ref) < 1e-5
    gy = torch.randn_like(y_ref)
    y.backward(gy); y_ref.backward(gy)
    assert x16.grad.dtype == torch.bfloat16 and rel_err(x16.grad, x32.grad) < 1e-2
    enc = _torch_chain((4, 16, 32, 64), False).to(DEV)
    a = torch.randn(M, 4, device=DEV)
    e16 = ops.run_mlp(enc, [(a, None)], out_dtype=torch.bfloat16)
    assert e16.dtype == torch.bfloat16 and rel_err(e16, enc(a)) < 1e-2
    # float64 input rows (edge_attr as stored): the `.float()` cast happens in the kernel's row load
    a64 = a.double()
    e64in = ops.run_mlp(enc, [(ops.edge_attr_rows(a64), None)], out_dtype=torch.bfloat16)
    assert torch.equal(e64in, e16)
    g64 = torch.autograd.grad(e64in, list(enc.parameters()), torch.ones_like(e64in))
    g32 = torch.autograd.grad(ops.run_mlp(enc, [(a, None)], out_dtype=torch.bfloat16), list(enc.parameters()),
                              torch.ones_like(e16))
    assert all(torch.equal(p, q) for p, q in zip(g64, g32))
    g16 = torch.randn(M, 64, device=DEV).to(torch.bfloat16)
    grads = torch.autograd.grad(e16, list(enc.parameters()), g16)
    grads_ref = torch.autograd.grad(enc(a), list(enc.parameters()), g16.float())
    for g, r in zip(grads, grads_ref):
        assert rel_err(g, r) < 1e-4


@pytest.mark.parametrize("M", [1, 255, 1000, 70001])
def test_narrow_mlp_partial_chains(M):
    """The narrow remainders of the bf16 mode's split chains (ops._narrow_split), fp32 arithmetic against torch
    autograd: ReLU -> 32 -> 16 -> 8 -> 1 -> Sigmoid (a ReLU on the chain INPUT, dX masked by it) and
    4 -> 16 -> 32 -> ReLU (a 2-layer chain ending in a ReLU)."""
    torch.manual_seed(M)
    for dims, in_relu, final in (((32, 16, 8, 1), True, "sigmoid"), ((32, 16, 8, 1), True, None), ((4, 16, 32), False, "relu"),
                                 ((4, 16, 32), False, None)):
        seq = _torch_chain(dims, final == "sigmoid").double()
        x = torch.randn(M, dims[0]).double().requires_grad_(True)
        h = torch.relu(x) if in_relu else x
        y_ref = seq(h)
        if final == "relu":
            y_ref = torch.relu(y_ref)
        gy = torch.randn_like(y_ref)
        y_ref.backward(gy)
        lin = [m for m in seq if isinstance(m, torch.nn.Linear)]
        params = []
        for m in lin:
            params += [m.weight.detach().float().to(DEV).requires_grad_(True), m.bias.detach().float().to(DEV).requires_grad_(True)]
        xd = x.detach().float().to(DEV).requires_grad_(True)
        y = ops._NarrowMLP.apply(len(lin), final, torch.float32, in_relu, xd, *params)
        assert rel_err(y, y_ref) < 1e-5
        y.backward(gy.float().to(DEV))
        assert rel_err(xd.grad, x.grad) < 1e-4
        for m, (w, b) in zip(lin, zip(params[0::2], params[1::2])):
            assert rel_err(w.grad, m.weight.grad) < 1e-4 and rel_err(b.grad, m.bias.grad) < 1e-4


def test_narrow_chains_split_in_bf16_mode():
    """bf16 mode: classifier 64 -> 32 and edge encoder 32 -> 64 run as tensor-core layers, the rest in the narrow
    kernel; outputs and gradients against torch fp32 on the same (bf16-representable) inputs within the bf16 tolerance."""
    torch.manual_seed(11)
    M = 20000
    ops.set_precision("bf16")
    try:
        for dims, sig in (((64, 32, 16, 8, 1), True), ((4, 16, 32, 64), False)):
            seq = _torch_chain(dims, sig).to(DEV)
            x = torch.randn(M, dims[0], device=DEV)
            x = x.to(torch.bfloat16) if dims[0] == 64 else x.double()
            xr = x.float().requires_grad_(True)
            y_ref = seq(xr)
            gy = torch.randn_like(y_ref)
            gref = torch.autograd.grad(y_ref, [xr] + list(seq.parameters()), gy)
            xin = x.clone().requires_grad_(dims[0] == 64)
            before = L.launch_count()
            y = ops.run_mlp(seq, [(ops.edge_attr_rows(xin) if dims[0] == 4 else xin, None)],
                            final_act="sigmoid" if sig else None, out_dtype=torch.float32 if sig else torch.bfloat16)
            assert L.launch_count() - before == 3, "pack + tensor-core layer + narrow kernel"
            assert rel_err(y, y_ref) < 2e-2
            got = torch.autograd.grad(y, ([xin] if dims[0] == 64 else []) + list(seq.parameters()), gy.to(y.dtype))
            # Frobenius norms, loose: rounding the 32-wide hidden row to bf16 moves ~0.25 % of the downstream units
            # across their ReLU threshold (24 units: ~5 % of the rows), which switches those rows' gradient paths; with
            # the RANDOM output gradient of this test nothing averages out (measured: dW of the tensor-core layer 5.2e-2,
            # dX similar; forward 2e-3). The model-level tests (coherent gradients) hold 2e-2 on the same kernels.
            for i, (g, r) in enumerate(zip(got, gref[0 if dims[0] == 64 else 1:])):
                assert float((g.double() - r.double()).norm() / r.double().norm()) < 0.12, (i, dims)
    finally:
        ops.set_precision("exact")
        ops.invalidate_weight_cache()
