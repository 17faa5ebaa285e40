"""Single-layer programs through k_chain (16-warp epilogue) vs k_linear_tma (8-warp epilogue), CUDA events."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("B3D_FEATURES", "all")
from batch3dmot_b200 import _lib as L, ops  # noqa: E402

dev = "cuda"
scenes = int(sys.argv[1]) if len(sys.argv) > 1 else 32
M, Nn = 61220 * scenes, 2000 * scenes
torch.manual_seed(0)
bf = torch.bfloat16
dst = torch.sort(torch.randint(0, Nn, (M,), device=dev, dtype=torch.int32)).values
src = (dst.long() - torch.randint(1, 2000, (M,), device=dev)).clamp(min=0).int()
rnd = lambda *s: (torch.randn(*s, device=dev) * 0.5).to(bf)
W = lambda n, k: torch.randn(n, k, device=dev) / k ** 0.5
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


ops.set_precision("bf16")
cases = [("64->192 relu +1 gathered add, bits", 64, 192, 1, True, False),
         ("128->256 relu +2 gathered adds, bits", 128, 256, 2, True, False),
         ("256->128 relu plain", 256, 128, 0, False, False),
         ("128->256 dgrad maskbits", 128, 256, 0, False, True),
         ("64->512 relu +2 adds", 64, 512, 2, False, False),
         ("512->384 relu plain", 512, 384, 0, False, False),
         ("384->256 relu plain", 384, 256, 0, False, False)]
for name, K, N, nadd, bits, maskbits in cases:
    x = rnd(M, K)
    w = W(N, K)
    adds = [rnd(Nn, N) for _ in range(nadd)]
    idx = [dst, src][:nadd]
    y = mk(N)
    bt = ops.new_relu_bits(M, N, dev) if bits else None
    mb = ops.new_relu_bits(M, N, dev).random_() if maskbits else None
    spec = [dict(W=w, src=-1, act=L.ACT_MASKBITS if maskbits else L.ACT_RELU, adds=[(a, t) for t, a in enumerate(adds)],
                 out=y, bits_out=bt, bits_in=mb)]
    ok = ops.chain_run([x], spec, dst, src, M)
    t_chain = timeit(lambda: ops.chain_run([x], spec, dst, src, M)) if ok else float("nan")
    items = [(x, None, None, 0)]
    y2 = mk(N)
    f = lambda: ops.linear_raw(items, w, None, M, 0 if maskbits else L.ACT_RELU, out=y2, tc=True, out_dtype=bf,
                               adds=[(a, i) for a, i in zip(adds, idx)] or None, bits_out=bt, mask_bits=mb)
    t_lin = timeit(f)
    nbytes = M * (K + N) * 2
    print(f"{name:40s} k_chain {t_chain:8.1f} us ({nbytes / t_chain / 1e3:6.0f} GB/s)   k_linear_tma {t_lin:8.1f} us ({nbytes / t_lin / 1e3:6.0f} GB/s)")
