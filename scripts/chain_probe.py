"""Times the fused chain programs against the per-layer kernels they replace (CUDA events), and is the target of
`ncu --set full -k regex:k_chain` captures:  python scripts/chain_probe.py [scenes]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("B3D_FEATURES", "all")
from batch3dmot_b200 import _lib as L, ops  # noqa: E402

dev = "cuda"
scenes = int(sys.argv[1]) if len(sys.argv) > 1 else 32
M, Nn = 61220 * scenes, 2000 * scenes
torch.manual_seed(0)
bf = torch.bfloat16
src = torch.randint(0, Nn, (M,), device=dev, dtype=torch.int32)
dst = torch.sort(torch.randint(0, Nn, (M,), device=dev, dtype=torch.int32)).values
# scene-local locality like the real graphs: sources within +-2000 of the target
src = (dst.long() - torch.randint(1, 2000, (M,), device=dev)).clamp(min=0).int()
rnd = lambda *s: (torch.randn(*s, device=dev) * 0.5).to(bf)
W = lambda n, k: torch.randn(n, k, device=dev) / k ** 0.5
b = lambda n: torch.randn(n, device=dev) * 0.1
mk = lambda n: torch.empty(M, n, dtype=bf, device=dev)


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


e, att = rnd(M, 64), rnd(M, 64)
p_i, p_j, p_f, p_p = rnd(Nn, 256), rnd(Nn, 256), rnd(Nn, 192), rnd(Nn, 192)
W0, W1, W2, Wf, Wp = W(256, 128), W(128, 256), W(64, 128), W(192, 64), W(192, 64)
b1, b2 = b(128), b(64)
R_, N_ = L.ACT_RELU, L.ACT_NONE
outs = dict(x1=mk(256), x2=mk(128), e2=mk(64), hf=mk(192), hp=mk(192))


def mp_specs(train, branch=True, adds=True):
    sp = [dict(W=W0, src=-1, act=R_, adds=[(p_i, 0), (p_j, 1)] if adds else None, out=outs["x1"] if train else None),
          dict(W=W1, src=0, act=R_, bias=b1, out=outs["x2"] if train else None),
          dict(W=W2, src=1, act=N_, bias=b2, out=outs["e2"])]
    if branch:
        sp += [dict(W=Wf, src=2, act=R_, adds=[(p_f, 0)] if adds else None, out=outs["hf"]),
               dict(W=Wp, src=2, act=R_, adds=[(p_p, 1)] if adds else None, out=outs["hp"])]
    return sp


res = {}
for name, sp in (("mp5_infer", mp_specs(False)), ("mp5_train", mp_specs(True)), ("mp3_infer", mp_specs(False, False)),
                 ("mp5_infer_noadds", mp_specs(False, True, False)), ("mp3_infer_noadds", mp_specs(False, False, False))):
    assert ops.chain_run([e, att], sp, dst, src, M)
    res[name] = timeit(lambda: ops.chain_run([e, att], sp, dst, src, M))
# att_edge_encoder head / tail
e0 = rnd(M, 64)
q_i, q_j = rnd(Nn, 512), rnd(Nn, 512)
A0, A1, A2, A3, A4 = W(512, 64), W(384, 512), W(256, 384), W(128, 256), W(64, 128)
x384, y64 = mk(384), mk(64)
head = [dict(W=A0, src=-1, act=R_, adds=[(q_i, 0), (q_j, 1)]), dict(W=A1, src=0, act=R_, bias=b(384), out=x384)]
tail = [dict(W=A2, src=-1, act=R_, bias=b(256)), dict(W=A3, src=0, act=R_, bias=b(128)), dict(W=A4, src=1, act=N_, bias=b(64), out=y64)]
assert ops.chain_run([e0], head, dst, src, M) and ops.chain_run([x384], tail, None, None, M)
res["att_head_infer"] = timeit(lambda: ops.chain_run([e0], head, dst, src, M))
res["att_tail_infer"] = timeit(lambda: ops.chain_run([x384], tail, None, None, M))

# the per-layer kernels they replace (bf16 mode, inference: no bits)
ops.set_precision("bf16")
ops._USE_CHAIN = False
g = type("G", (), {})()
from batch3dmot_b200.ops import NodeIndex  # noqa: E402
nd, ns = NodeIndex(dst, None, None, Nn, True), NodeIndex(src, None, None, Nn)
with torch.no_grad():
    f_eu = lambda: ops.fused_mlp([(e, None), (att, None)], [W0, W1, W2], [None, b1, b2], adds=[(p_i, nd), (p_j, ns)], out_dtype=bf)
    res["per_layer_edge_update"] = timeit(f_eu)
    e2 = outs["e2"]
    f_m = lambda: (ops.fused_mlp([(e2, None)], [Wf], [None], final_act="relu", adds=[(p_f, nd)], out_dtype=bf, premasked=True),
                   ops.fused_mlp([(e2, None)], [Wp], [None], final_act="relu", adds=[(p_p, ns)], out_dtype=bf, premasked=True))
    res["per_layer_two_message_layers"] = timeit(f_m)
    f_att = lambda: ops.fused_mlp([(e0, None)], [A0, A1, A2, A3, A4], [None, b(384), b(256), b(128), b(64)], adds=[(q_i, nd), (q_j, ns)], out_dtype=bf)
    res["per_layer_att_edge_encoder"] = timeit(f_att)
tiles = (M + 127) // 128
print(f"E = {M} rows, {tiles} tiles, {tiles / 148:.0f} tiles per SM")
for k, v in res.items():
    print(f"  {k:32s} {v:9.1f} us   {v / (tiles / 148):6.2f} us per tile-wave")
