"""GPU parity of the fused MLP-chain kernel (chain_tc.cu, b3d_chain_run): every layer output, every sign-bit mask
and the backward (input-gradient) chains against a float64 reference that rounds to bf16 exactly where the kernel
does (operands, hidden activations); only the fp32 accumulation order inside a tcgen05 tile differs."""
import pytest
import torch

from batch3dmot_b200 import _lib as L, ops

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(autouse=True)
def chain_on():
    """The fused kernels are exercised here whatever the model-path default (ops.FEATURES["chain"]) is."""
    old = ops._USE_CHAIN
    ops._USE_CHAIN = True
    yield
    ops._USE_CHAIN = old


def bf(t):
    return t.to(torch.bfloat16).to(torch.float64)


def unpack_bits(bits, n):
    """uint32 words [ceil(n/32), M] -> bool [M, n]"""
    w = bits.cpu().to(torch.int64) & 0xFFFFFFFF
    cols = torch.arange(n)
    return ((w[cols // 32] >> (cols % 32)[:, None]) & 1).bool().t()


def check(got, ref, what, tol=1e-2):
    ref = ref.double()
    err = float((got.double().cpu() - ref).abs().max() / ref.abs().max().clamp_min(1e-30))
    assert err < tol, (what, err)


def run_program(M, in_widths, layers, n_nodes=500, seed=0):
    """layers: list of (K, N, src, act, n_adds, has_bias). Returns nothing; asserts parity of every output."""
    g = torch.Generator().manual_seed(seed)
    xs = [torch.randn(M, w, generator=g).to(torch.bfloat16) for w in in_widths]
    idx = [torch.randint(0, n_nodes, (M,), generator=g, dtype=torch.int32) for _ in range(2)]
    x_cat = torch.cat([x.double() for x in xs], 1)
    specs, refs, ref_bits = [], [], []
    for l, (K, N, src, act, nadd, has_bias) in enumerate(layers):
        W = torch.randn(N, K, generator=g) * (1.0 / K ** 0.5)
        b = torch.randn(N, generator=g) * 0.1 if has_bias else None
        adds = [(torch.randn(n_nodes, N, generator=g) * 0.5).to(torch.bfloat16) for _ in range(nadd)]
        a_in = x_cat if src < 0 else refs[src]
        z = a_in @ bf(W).t()
        if b is not None:
            z = z + b.double()
        for t, ad in enumerate(adds):
            z = z + ad.double()[idx[t].long()]
        bits_in = None
        if act == L.ACT_RELU:
            z = torch.relu(z)
        elif act == L.ACT_MASKBITS:
            mask = torch.rand(M, N, generator=g) < 0.5
            words = torch.zeros((N // 32, M), dtype=torch.int64)
            for c in range(N):
                words[c // 32] |= mask[:, c].long() << (c % 32)
            bits_in = torch.where(words >= 2 ** 31, words - 2 ** 32, words).to(torch.int32).to(DEV)
            z = z * mask
        refs.append(bf(z))
        ref_bits.append(z > 0)
        specs.append(dict(W=W.to(DEV), src=src, act=act, bias=b.to(DEV) if b is not None else None,
                          adds=[(ad.to(DEV), t) for t, ad in enumerate(adds)],
                          out=torch.empty(M, N, dtype=torch.bfloat16, device=DEV),
                          bits_out=ops.new_relu_bits(M, N, DEV) if act == L.ACT_RELU else None, bits_in=bits_in))
    n0 = L.launch_count()
    ok = ops.chain_run([x.to(DEV) for x in xs], specs, idx[0].to(DEV), idx[1].to(DEV), M)
    assert ok, "chain plan rejected"
    assert L.launch_count() - n0 == len(layers) + 1          # pack per layer + ONE fused launch
    torch.cuda.synchronize()
    for l, sp in enumerate(specs):
        check(sp["out"], refs[l], f"layer {l} output")
        if sp["bits_out"] is not None:
            got = unpack_bits(sp["bits_out"], layers[l][1])
            # a sign can only differ where the pre-activation is within rounding noise of zero
            bad = (got != ref_bits[l]) & (refs[l].abs() > 1e-2)
            assert int(bad.sum()) == 0, (l, int(bad.sum()))
            assert float((got != ref_bits[l]).float().mean()) < 2e-3
    # second run hits the pack cache: one launch
    n0 = L.launch_count()
    ops.chain_run([x.to(DEV) for x in xs], specs, idx[0].to(DEV), idx[1].to(DEV), M)
    assert L.launch_count() - n0 == 1


R, N_, MB = L.ACT_RELU, L.ACT_NONE, L.ACT_MASKBITS


@pytest.mark.parametrize("M", [128, 1000, 148 * 128 * 2 + 77])
def test_chain_message_passing_program(M):
    """edge_update cat[e, att] -> 256 -> 128 -> 64 with two gathered node-side addends, then both message first layers
    reading e' (a two-way branch), each with its own gathered addend."""
    run_program(M, (64, 64), [(128, 256, -1, R, 2, True), (256, 128, 0, R, 0, True), (128, 64, 1, N_, 0, True),
                              (64, 192, 2, R, 1, False), (64, 192, 2, R, 1, False)], seed=M)


@pytest.mark.parametrize("M", [300, 40000])
def test_chain_att_edge_encoder_head(M):
    """att_edge_encoder layers 0-1: 64 (+2 addends) -> 512 -> 384 (the 512-wide activation tile is 128 KB)."""
    run_program(M, (64,), [(64, 512, -1, R, 2, True), (512, 384, 0, R, 0, True)], seed=M)


def test_chain_att_edge_encoder_tail():
    run_program(5000, (384,), [(384, 256, -1, R, 0, True), (256, 128, 0, R, 0, True), (128, 64, 1, N_, 0, True)], seed=3)


def test_chain_backward_program():
    """dZ chain of edge_update: 64 -> 128 (mask) -> 256 (mask) -> 128 (no mask), no bias."""
    run_program(7000, (64,), [(64, 128, -1, MB, 0, False), (128, 256, 0, MB, 0, False), (256, 128, 1, N_, 0, False)], seed=5)


def test_chain_rejects_what_it_cannot_plan():
    x = torch.zeros(256, 64, dtype=torch.bfloat16, device=DEV)
    W = torch.zeros(48, 64, device=DEV)                     # N % 64 != 0
    assert ops.chain_run([x], [dict(W=W, out=torch.empty(256, 48, dtype=torch.bfloat16, device=DEV))], None, None, 256) is False


def test_fused_mlp_uses_the_chain_and_matches_the_per_layer_path():
    """ops.fused_mlp in bf16 mode: fused launch == per-layer launches up to one bf16 rounding step (same rounding
    points; the per-layer kernel adds the gathered addends inside the tensor-core accumulator, the fused kernel in the
    epilogue, so the fp32 pre-activations can differ in the last bit), forward and backward (input gradient, weight
    gradients, addend gradients)."""
    torch.manual_seed(0)
    M, Nn = 6000, 300
    ei = torch.randint(0, Nn, (2, M))
    ei = ei[:, torch.argsort(ei[1], stable=True)].to(DEV)
    g = ops.Graph(ei, Nn)
    e = torch.randn(M, 64, device=DEV).to(torch.bfloat16).requires_grad_(True)
    att = torch.randn(M, 64, device=DEV).to(torch.bfloat16).requires_grad_(True)
    p_i = torch.randn(Nn, 256, device=DEV).to(torch.bfloat16).requires_grad_(True)
    p_j = torch.randn(Nn, 256, device=DEV).to(torch.bfloat16).requires_grad_(True)
    Ws = [torch.nn.Parameter(torch.randn(n, k, device=DEV) / k ** 0.5) for k, n in ((128, 256), (256, 128), (128, 64))]
    bs = [None] + [torch.nn.Parameter(torch.randn(n, device=DEV) * 0.1) for n in (128, 64)]
    dy = torch.randn(M, 64, device=DEV).to(torch.bfloat16)
    ops.set_precision("bf16")
    try:
        res = {}
        for fused in (True, False):
            ops._USE_CHAIN = fused
            for t in [e, att, p_i, p_j] + Ws + bs[1:]:
                t.grad = None
            n0 = L.launch_count()
            y = ops.fused_mlp([(e, None), (att, None)], Ws, bs, adds=[(p_i, g.by_dst), (p_j, g.by_src)],
                              out_dtype=torch.bfloat16)
            fwd_launches = L.launch_count() - n0
            y.backward(dy)
            res[fused] = (y.detach().clone(), [t.grad.clone() for t in [e, att, p_i, p_j] + Ws + bs[1:]], fwd_launches)
        assert res[True][2] < res[False][2]
        # forward: last-bit differences only. Gradients: a pre-activation within rounding noise of zero can land on
        # either side of the ReLU in the two paths, which switches single gradient elements on or off: compare norms.
        y1, y0 = res[True][0].float(), res[False][0].float()
        assert float((y1 - y0).abs().max()) <= 2 ** -7 * float(y0.abs().max())
        assert float((y1 != y0).float().mean()) < 2e-2
        for a, b in zip(res[True][1], res[False][1]):
            a, b = a.double(), b.double()
            assert float((a - b).norm() / b.norm().clamp_min(1e-30)) < 1e-2
    finally:
        ops.set_precision("fp32")
        ops.invalidate_weight_cache()
