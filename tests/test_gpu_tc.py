"""GPU parity of the bf16 tensor-core mode (tcgen05 tiles, fp32 accumulation in TMEM).
Kernel-level: against a float64 reference on bf16-rounded operands (tight: only the accumulation
order differs). Model-level: against the fp32 CPU oracle within the 2e-2 tolerance north_star
states for bf16 MLP tiles."""
from types import SimpleNamespace

import pytest
import torch

from oracle import ref_restated as R
from batch3dmot_b200 import _lib as L, ops, synth
from batch3dmot_b200.pose_gnn import PoseGNN
from batch3dmot_b200.clr_att_gnn import GNN

pytestmark = pytest.mark.gpu
DEV = "cuda"
BF16_TOL = 2e-2


@pytest.fixture(autouse=True)
def bf16_mode():
    ops.set_precision("bf16")
    ops.invalidate_weight_cache()
    yield
    ops.set_precision("fp32")
    ops.invalidate_weight_cache()


def bf(t):
    return t.to(torch.bfloat16).to(torch.float64)


def rel(a, b):
    b = b.double().cpu()
    return float((a.detach().double().cpu() - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _operands(M, widths, N=777):
    xs = [torch.randn(N if s % 2 == 0 else M, w) for s, w in enumerate(widths)]
    idx = torch.randint(0, N, (len(widths), M))
    cat = torch.cat([x[idx[s]] if x.size(0) == N else x for s, x in enumerate(xs)], 1)
    items = [(x.to(DEV), idx[s].int().to(DEV) if x.size(0) == N else None, None, 0) for s, x in enumerate(xs)]
    return cat, items


@pytest.mark.parametrize("M,widths,n_out", [(1024, (64,), 64), (3000, (48, 48, 32), 96), (5000, (96, 96, 64, 64), 256),
                                            (2000, (64, 128, 96, 64, 128, 96, 64), 512), (1500, (256,), 128),
                                            (700, (128,), 48), (300, (192,), 128), (129, (32,), 16)])
def test_linear_tc(M, widths, n_out):
    torch.manual_seed(M + n_out)
    cat, items = _operands(M, widths)
    W, b = torch.randn(n_out, sum(widths)) * 0.1, torch.randn(n_out)
    launches = L.launch_count()
    y = ops.linear_raw(items, W.to(DEV), b.to(DEV), M, L.ACT_RELU)
    assert M < 256 or L.launch_count() - launches == 2          # pack + k_linear_tc really ran
    ref0 = torch.relu(bf(cat) @ bf(W).t() + b.double())
    assert rel(y, ref0) < (1e-5 if M >= 256 else 1e-2)
    # input gradient form (transposed pack): ReLU mask of the producing layer applied in the epilogue
    # (fp32 or bf16 saved activation), bf16 or fp32 gradient operand, bf16 or fp32 output
    if n_out % 8 == 0 and M >= 256 and n_out >= 32:
        dy, hmask = torch.randn(M, n_out), torch.randn(M, sum(widths))
        ref = (bf(dy) @ bf(W)) * (hmask > 0)
        for dy_dt, m_dt, o_dt in ((torch.float32, torch.float32, torch.float32),
                                  (torch.bfloat16, torch.bfloat16, torch.bfloat16),
                                  (torch.float32, torch.bfloat16, torch.float32)):
            out = ops.linear_raw([(dy.to(DEV).to(dy_dt), None, None, 0)], W.to(DEV), None, M, trans_w=True,
                                 out_mask=hmask.to(DEV).to(m_dt), tc=True, out_dtype=o_dt)
            assert out.dtype == o_dt
            assert rel(out, ref) < (1e-5 if o_dt == torch.float32 else 1e-2)
    # bf16 input segments (cp.async staging) and bf16 output
    if all(w % 8 == 0 for w in widths) and M >= 256 and n_out >= 16 and sum(widths) >= 32:
        items16 = [(t.to(torch.bfloat16), i, None, 0) for t, i, _, _ in items]
        y16 = ops.linear_raw(items16, W.to(DEV), b.to(DEV), M, L.ACT_RELU, tc=True, out_dtype=torch.bfloat16)
        assert y16.dtype == torch.bfloat16 and rel(y16, ref0) < 1e-2


@pytest.mark.parametrize("M,widths,n_out", [(4096, (64,), 64), (3000, (48, 48, 32), 96), (70000, (96, 96, 64, 64), 256),
                                            (9000, (96, 64, 96), 192), (5000, (128,), 64), (20000, (640,), 512),
                                            (777, (64,), 32)])
def test_wgrad_tc(M, widths, n_out):
    torch.manual_seed(M)
    cat, items = _operands(M, widths, N=333)
    dy = torch.randn(M, n_out)
    ref_w, ref_b = bf(dy).t() @ bf(cat), bf(dy).sum(0)
    dW, db = ops.wgrad_raw((dy.to(DEV), None, None, 0), items, M, n_out, sum(widths), tc=True)
    assert rel(dW, ref_w) < 1e-5 and rel(db, ref_b) < 1e-5
    dW2, _ = ops.wgrad_raw((dy.to(DEV), None, None, 0), items, M, n_out, sum(widths), tc=True)
    assert torch.equal(dW, dW2)                                   # deterministic
    dW3, db3 = ops.wgrad_raw((dy.to(DEV), None, None, 0), items, M, n_out, sum(widths),
                             dW=dW.clone(), db=db.clone(), accumulate=True, tc=True)
    assert rel(dW3, 2 * dW) < 1e-6 and rel(db3, 2 * db) < 1e-6
    # bf16 operands (cp.async staging path)
    items16 = [(t.to(torch.bfloat16), i, None, 0) for t, i, _, _ in items]
    dW4, db4 = ops.wgrad_raw((dy.to(DEV).to(torch.bfloat16), None, None, 0), items16, M, n_out, sum(widths), tc=True)
    assert rel(dW4, ref_w) < 1e-5 and rel(db4, ref_b) < 1e-5


def to_dev(ns):
    return SimpleNamespace(**{k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in vars(ns).items()})


def _grad_check(model, gref, tol):
    """bf16 rounding noise is unstructured, so gradients are compared in the Frobenius norm
    (||got - ref|| / ||ref|| < tol) plus a loose max-entry bound."""
    for k, p in model.named_parameters():
        if k.startswith("knn_conv") or k not in gref:
            continue
        got, ref = p.grad.detach().double().cpu(), gref[k].double()
        if "in_proj" in k:
            D = ref.size(0) // 3
            got, ref = got[2 * D:], ref[2 * D:]
        fro = float((got - ref).norm() / ref.norm().clamp_min(1e-30))
        assert fro < tol, (k, fro)
        assert rel(got, ref) < 6 * tol, (k, rel(got, ref))


def test_pose_gnn_bf16_vs_oracle():
    data = synth.add_labels(synth.scene_graph(seed=5621), 5621)
    torch.manual_seed(5621)
    m = PoseGNN()
    params = {k: v.clone().requires_grad_(True) for k, v in m.state_dict().items()}
    out_ref, _ = R.pose_gnn_forward(params, data)
    loss_ref = R.bce_logits_loss(out_ref, data.y, data.edge_weights)
    loss_ref.backward()
    m = m.to(DEV)
    d = to_dev(data)
    launches = L.launch_count()
    out, _ = m(d)
    loss = ops.bce_loss(out, d.y, d.edge_weights, from_logits=True)
    loss.backward()
    assert L.launch_count() > launches
    assert rel(out, out_ref) < BF16_TOL
    assert abs(loss.item() - loss_ref.item()) < BF16_TOL * abs(loss_ref.item())
    _grad_check(m, {k: v.grad for k, v in params.items() if v.grad is not None}, 8e-2)


def test_mm_gnn_bf16_vs_oracle():
    data = synth.add_labels(synth.add_modalities(synth.scene_graph(seed=5621), 5621, raw=False), 5621)
    torch.manual_seed(5621)
    m = GNN(None, None, None)
    params = {k: v.clone().requires_grad_(True) for k, v in m.state_dict().items()}
    out_ref, _ = R.mm_gnn_forward(params, data)
    loss_ref = R.bce_loss(out_ref, data.y, data.edge_weights, batch_size=2)
    loss_ref.backward()
    m = m.to(DEV)
    d = to_dev(data)
    out, x_sens = m(d, x_img=d.x_img, pointnet_out=d.pointnet_out, radarnet_out=d.radarnet_out, lidar_mask=d.m_lidar,
                    radar_mask=d.m_radar)
    loss = ops.bce_loss(out, d.y, d.edge_weights, batch_size=2)
    loss.backward()
    assert rel(out, out_ref) < BF16_TOL
    assert abs(loss.item() - loss_ref.item()) < BF16_TOL * abs(loss_ref.item())
    _grad_check(m, {k: v.grad for k, v in params.items() if v.grad is not None}, 8e-2)


@pytest.mark.parametrize("M,widths,n_out", [(1024, (64,), 64), (5000, (64, 64), 256), (70000, (256,), 128),
                                            (3000, (192,), 128), (2000, (512,), 384), (9000, (384,), 256),
                                            (4000, (128, 40), 96), (130, (128,), 64), (100000, (64,), 512)])
def test_linear_tma_dense_bf16(M, widths, n_out):
    """TMA-fed persistent kernel (dense bf16 operands): all epilogue variants vs a float64 reference."""
    torch.manual_seed(M + n_out)
    xs = [torch.randn(M, w).to(torch.bfloat16) for w in widths]
    cat = torch.cat([x.double() for x in xs], 1)
    items = [(x.to(DEV), None, None, 0) for x in xs]
    W, b = torch.randn(n_out, sum(widths)) * 0.1, torch.randn(n_out)
    ref = bf(cat) @ bf(W).t() + b.double()
    before = L.launch_count()
    y = ops.linear_raw(items, W.to(DEV), b.to(DEV), M, L.ACT_RELU, tc=True, out_dtype=torch.float32)
    assert L.launch_count() - before == 2          # row-major pack + k_linear_tma
    assert rel(y, torch.relu(ref)) < 1e-5
    y16 = ops.linear_raw(items, W.to(DEV), b.to(DEV), M, L.ACT_NONE, tc=True, out_dtype=torch.bfloat16)
    assert y16.dtype == torch.bfloat16 and rel(y16, ref) < 1e-2
    # row-gathered addends + ReLU-mask epilogue + row mask
    N = 501
    p0, p1 = torch.randn(N, n_out), torch.randn(N, n_out)
    i0, i1 = torch.randint(0, N, (M,)), torch.randint(0, N, (M,))
    hm = torch.randn(M, n_out)
    rm = torch.rand(M) < 0.8
    y = ops.linear_raw(items, W.to(DEV), b.to(DEV), M, L.ACT_NONE, tc=True, out_mask=hm.to(DEV).to(torch.bfloat16),
                       row_mask=rm.to(torch.uint8).to(DEV),
                       adds=[(p0.to(DEV), i0.int().to(DEV)), (p1.to(DEV), i1.int().to(DEV))])
    ref2 = (ref + p0[i0].double() + p1[i1].double()) * (bf(hm) > 0) * rm[:, None]
    assert rel(y, ref2) < 1e-5
    # bf16 addends / bf16 mask / bf16 output: the operands the pipelined epilogue prefetches ahead of
    # the accumulator (mask only, one addend, two addends, mask + addends)
    p0h, p1h = p0.to(torch.bfloat16), p1.to(torch.bfloat16)
    a0, a1 = (p0h.to(DEV), i0.int().to(DEV)), (p1h.to(DEV), i1.int().to(DEV))
    hm16 = hm.to(DEV).to(torch.bfloat16)
    for adds, mask, act in (([], hm16, L.ACT_NONE), ([a0], None, L.ACT_RELU), ([a0, a1], None, L.ACT_RELU),
                            ([a0, a1], hm16, L.ACT_NONE), ([(p0h[i0].to(DEV), None)], None, L.ACT_NONE)):
        r3 = ref.clone()
        for (t, ix) in adds:
            r3 = r3 + (t.cpu().double()[ix.cpu().long()] if ix is not None else t.cpu().double())
        if act == L.ACT_RELU:
            r3 = torch.relu(r3)
        if mask is not None:
            r3 = r3 * (bf(hm) > 0)
        for od in (torch.bfloat16, torch.float32):
            y3 = ops.linear_raw(items, W.to(DEV), b.to(DEV), M, act, tc=True, out_mask=mask, adds=adds or None,
                                out_dtype=od)
            assert y3.dtype == od and rel(y3, r3) < (1e-2 if od == torch.bfloat16 else 1e-5)
    # ReLU sign bits: written by the forward epilogue (B3D_BITS layout), then used as the mask of an
    # input-gradient launch; both kernels (TMA-fed and thread-staged) must agree with the bf16 mask
    if n_out % 32 == 0:
        for use_tma in (True, False):
            ops._USE_TMA = use_tma
            try:
                bits = ops.new_relu_bits(M, n_out, DEV)
                yb = ops.linear_raw(items, W.to(DEV), b.to(DEV), M, L.ACT_RELU, tc=True, out_dtype=torch.bfloat16,
                                    bits_out=bits)
                unpacked = ((bits.t().unsqueeze(2) >> torch.arange(32, device=DEV, dtype=torch.int32)) & 1)
                unpacked = unpacked.reshape(M, -1)[:, :n_out].bool()
                assert torch.equal(unpacked, yb > 0)
                # gradient w.r.t. a [M, n_out] ReLU output masked by its sign bits vs by the activation itself
                Wd = torch.randn(96, n_out, device=DEV) * 0.1          # consumer Linear(n_out -> 96)
                gz = torch.randn(M, 96).to(torch.bfloat16).to(DEV)
                d1 = ops.linear_raw([(gz, None, None, 0)], Wd, None, M, trans_w=True, tc=True, mask_bits=bits,
                                    out_dtype=torch.bfloat16)
                d2 = ops.linear_raw([(gz, None, None, 0)], Wd, None, M, trans_w=True, tc=True, out_mask=yb,
                                    out_dtype=torch.bfloat16)
                assert torch.equal(d1, d2)
            finally:
                ops._USE_TMA = True
    # same result as the thread-staged tcgen05 kernel
    ops._USE_TMA = False
    try:
        y_tc = ops.linear_raw(items, W.to(DEV), b.to(DEV), M, L.ACT_RELU, tc=True, out_dtype=torch.float32)
    finally:
        ops._USE_TMA = True
    y_tma = ops.linear_raw(items, W.to(DEV), b.to(DEV), M, L.ACT_RELU, tc=True, out_dtype=torch.float32)
    assert rel(y_tma, y_tc) < 1e-6


def test_sign_bits_with_column_blocks_past_the_last_word():
    """Regression: N = 512 with K = 384 plans three 192-column blocks (576 > 512); the blocks past the
    last 32-column word must neither read nor write sign bits (illegal address at full size)."""
    torch.manual_seed(5)
    M, n, k = 150000, 512, 384
    h = torch.randn(M, n, device=DEV).to(torch.bfloat16)
    bits = ops.new_relu_bits(M, n, DEV)
    w = (h > 0).reshape(M, n // 32, 32).to(torch.int64) << torch.arange(32, device=DEV)
    bits.copy_(w.sum(2).t().to(torch.int32))          # low 32 bits, two's complement
    gz = torch.randn(M, k, device=DEV).to(torch.bfloat16)
    W = torch.randn(k, n, device=DEV) * 0.1
    d1 = ops.linear_raw([(gz, None, None, 0)], W, None, M, trans_w=True, tc=True, mask_bits=bits, out_dtype=torch.bfloat16)
    d2 = ops.linear_raw([(gz, None, None, 0)], W, None, M, trans_w=True, tc=True, out_mask=h, out_dtype=torch.bfloat16)
    assert torch.equal(d1, d2)
    x = torch.randn(M, k, device=DEV).to(torch.bfloat16)
    b2 = ops.new_relu_bits(M, n, DEV)
    y = ops.linear_raw([(x, None, None, 0)], W.t().contiguous(), None, M, L.ACT_RELU, tc=True, out_dtype=torch.bfloat16,
                       bits_out=b2)
    un = ((b2.t().unsqueeze(2) >> torch.arange(32, device=DEV, dtype=torch.int32)) & 1).reshape(M, n).bool()
    assert torch.equal(un, y > 0)


@pytest.mark.parametrize("k,n,trans", [(288, 512, False), (512, 384, True), (288, 512, True), (320, 176, False)])
def test_linear_tma_column_blocks_not_multiple_of_32(k, n, trans):
    """Regression: with several column blocks per row tile whose width is 16 (mod 32), a CTA must not
    write past its own block (n = output features; trans = input-gradient form)."""
    torch.manual_seed(k + n)
    M = 20000
    x = torch.randn(M, n if trans else k).to(torch.bfloat16)
    W = torch.randn(n, k) * 0.1
    ref = bf(x) @ (bf(W) if trans else bf(W).t())
    for _ in range(3):      # repeated launches expose write races between neighbouring CTAs
        y = ops.linear_raw([(x.to(DEV), None, None, 0)], W.to(DEV), None, M, 0, trans_w=trans, tc=True)
        assert rel(y, ref) < 1e-5


@pytest.mark.parametrize("M,widths,n_out", [(4096, (64,), 64), (70000, (64, 64), 256), (9000, (256,), 128),
                                            (20000, (512,), 384), (5000, (192,), 128), (3000, (128, 40), 96),
                                            (1500, (64,), 512), (100000, (128,), 64), (777, (320,), 24)])
def test_wgrad_tma_dense_bf16(M, widths, n_out):
    """TMA-fed weight gradient (dense bf16 operands) incl. the ones-tile bias gradient."""
    torch.manual_seed(M + n_out)
    xs = [torch.randn(M, w).to(torch.bfloat16) for w in widths]
    cat = torch.cat([x.double() for x in xs], 1)
    dy = torch.randn(M, n_out).to(torch.bfloat16)
    items = [(x.to(DEV), None, None, 0) for x in xs]
    K = sum(widths)
    before = L.launch_count()
    dW, db = ops.wgrad_raw((dy.to(DEV), None, None, 0), items, M, n_out, K, tc=True)
    assert L.launch_count() - before == 2
    ref_w, ref_b = dy.double().t() @ cat, dy.double().sum(0)
    assert rel(dW, ref_w) < 1e-5 and rel(db, ref_b) < 1e-5
    dW2, db2 = ops.wgrad_raw((dy.to(DEV), None, None, 0), items, M, n_out, K, dW=dW.clone(), db=db.clone(),
                             accumulate=True, tc=True)
    assert rel(dW2, 2 * ref_w) < 1e-5 and rel(db2, 2 * ref_b) < 1e-5
    dW3, _ = ops.wgrad_raw((dy.to(DEV), None, None, 0), items, M, n_out, K, tc=True)
    assert torch.equal(dW3, dW)


@pytest.mark.parametrize("M,K,n_out", [(5000, 16, 8), (5000, 8, 16), (70000, 64, 32), (3000, 32, 16), (4000, 16, 32)])
def test_small_layers_on_tensor_core_tiles(M, K, n_out):
    """Narrow encoder / classifier layers (fp32 activations) through the thread-staged tcgen05 kernels."""
    torch.manual_seed(M + K)
    x, W, b = torch.randn(M, K), torch.randn(n_out, K) * 0.3, torch.randn(n_out)
    before = L.launch_count()
    y = ops.linear_raw([(x.to(DEV), None, None, 0)], W.to(DEV), b.to(DEV), M, L.ACT_RELU)
    assert L.launch_count() - before == 2 and y.dtype == torch.float32
    assert rel(y, torch.relu(bf(x) @ bf(W).t() + b.double())) < 1e-5
    dy = torch.randn(M, n_out)
    dW, db = ops.wgrad_raw((dy.to(DEV), None, None, 0), [(x.to(DEV), None, None, 0)], M, n_out, K)
    assert rel(dW, bf(dy).t() @ bf(x)) < 1e-5 and rel(db, bf(dy).sum(0)) < 1e-5
    hm = torch.randn(M, K)
    dx = ops.linear_raw([(dy.to(DEV), None, None, 0)], W.to(DEV), None, M, trans_w=True, out_mask=hm.to(DEV))
    assert rel(dx, (bf(dy) @ bf(W)) * (hm > 0)) < 1e-5
