"""CPU oracle for the Batch3DMOT tracking-graph GNN hot path.

TEST INFRASTRUCTURE ONLY. Nothing in `batch3dmot_b200/` (the product) may import
this package. Allowed importers: `tests/`, `__graft_entry__.smoke()` (as the
checker) and `bench.py`'s `cpu_baseline` / `--impl reference` legs.

Parity pin: the reference ships no tests or golden vectors, and
PyG / torch_scatter / torch_cluster are not installable here, so parity against
the *upstream third-party ops* is "parity unpinned". What IS pinned: the
reference's own unmodified model files (`/root/reference/batch_3dmot/models/
pose_gnn.py`, `clr_att_gnn.py`), executed in this container under
`oracle/pyg_shim.py`, agree bit-for-bit (max |diff| = 0.0 on CPU) with the
independent restatement in `oracle/ref_restated.py`, and the outputs of that
run are committed as fixtures under `tests/golden/` by `oracle/gen_golden.py`.
"""
