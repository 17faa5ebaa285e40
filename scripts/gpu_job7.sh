#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_split.py tests/test_round2_fixes.py tests/test_tracking.py tests/test_gpu_real_encoders.py -q -m gpu --timeout 300 --timeout-method=thread --tb=short > gpurun_out/r2_t7.log 2>&1
tail -15 gpurun_out/r2_t7.log | cut -c1-300
timeout 400 python - <<'PY'
import torch, bench, time
from batch3dmot_b200 import ops
from batch3dmot_b200.clr_att_gnn import GNN
from batch3dmot_b200.parallel import Trainer
dev = torch.device("cuda")
small = bench.to_dev(bench.make_batch(0, 8), dev)
small._b3d_graph = ops.Graph(small.edge_index, small.num_nodes)
E = small.edge_index.size(1)
tm = bench.Timer(dev, 1)
for mode in ("fp32", "exact"):
    ops.set_precision(mode)
    torch.manual_seed(5621)
    m = GNN(None, None, None).to(dev)
    tr = Trainer(m, batch_size=2)
    ms = tm.run(lambda: tr.step(small, **bench.mm_kwargs(small)), 3, 2, reduce=False)
    with torch.no_grad():
        msf = tm.run(lambda: m(small, **bench.mm_kwargs(small)), 3, 1, reduce=False)
    print(mode, "fwd+bwd edges/s %.2fM" % (E / ms / 1e3), "fwd %.2fM" % (E / msf / 1e3))
PY
