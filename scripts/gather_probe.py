"""Row-gathered TMA operands vs epilogue addends, per layer (CUDA events)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["B3D_FEATURES"] = "split_tc,window_knn,gather_tma"
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
it = lambda t, i=None: (t, i, None, 0)


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
x, x0, e, att = rnd(Nn, 96), rnd(Nn, 96), rnd(M, 64), rnd(M, 64)
x128 = rnd(Nn, 128)
cases = {
    "L0 gathered x_i|x_j|e|att (K=320->384) -> 256": ([it(x, dst), it(x, src), it(e), it(att)], 256, None),
    "L0 gathered x_j only |e|att (K=224->256) -> 256 + add p_i[dst]": ([it(e), it(att), it(x, src)], 256, [(rnd(Nn, 256), dst)]),
    "L0 pre-projected e|att -> 256 + 2 adds": ([it(e), it(att)], 256, [(rnd(Nn, 256), dst), (rnd(Nn, 256), src)]),
    "L0 dense only e|att -> 256 (no adds)": ([it(e), it(att)], 256, None),
    "msg gathered x(src)|e|x0(src) (K=256->320) -> 192": ([it(x, src), it(e), it(x0, src)], 192, None),
    "msg gathered x(dst)|e|x0(dst) -> 192": ([it(x, dst), it(e), it(x0, dst)], 192, None),
    "msg gathered one 128-wide table [x|x0pad](src)|e -> 192": ([it(x128, src), it(e)], 192, None),
    "msg pre-projected e -> 192 + add[src]": ([it(e)], 192, [(rnd(Nn, 192), src)]),
    "msg dense e -> 192 (no add)": ([it(e)], 192, None),
}
for name, (items, n, adds) in cases.items():
    K = sum(t.size(1) for t, _, _, _ in items)
    w = W(n, K)
    y = torch.empty(M, n, dtype=bf, device=dev)
    f = lambda: ops.linear_raw(items, w, None, M, L.ACT_RELU, out=y, tc=True, out_dtype=bf, adds=adds)
    print(f"{name:70s} {timeit(f):8.1f} us")
