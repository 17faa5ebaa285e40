"""TEST INFRASTRUCTURE — recipe that makes the UNMODIFIED reference model files available to the GPU box.

    python -m oracle.make_ref            (also run by __graft_entry__.build() in the build container)

`/root/reference` does not exist on the GPU box, and the reference cannot be pip-installed (pure-Python
research release without packaging metadata; its PyG / torch_scatter / torch_cluster dependencies are not
installable offline). What CAN travel is the reference's own model package, byte for byte: this script copies
`/root/reference/batch_3dmot/models/*.py` (pose_gnn.py, clr_att_gnn.py and the three encoder modules the
latter imports) into the git-ignored `oracle/_ref/batch_3dmot/models/` — outputs only, never committed, the
same way a compiled `oracle/_ref/*.so` of a C reference would travel. `oracle.pyg_shim.load_reference(root)`
imports them under the PyG stand-ins; `bench.py --impl reference` then times the reference's own
forward + BCELoss + backward + torch.optim.Adam step on the host cores, and falls back to the restated port
(`oracle/ref_restated.py`, pinned bit-exact to these files by tests/test_oracle.py) only when `_ref` is absent."""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/batch_3dmot/models"
DST = os.path.join(HERE, "_ref", "batch_3dmot", "models")


def ref_root():
    """Root to hand to pyg_shim.load_reference, or None when the copy was never made."""
    return os.path.join(HERE, "_ref") if os.path.exists(os.path.join(DST, "clr_att_gnn.py")) else None


def make(verbose=False):
    if not os.path.isdir(SRC):
        return ref_root()
    os.makedirs(DST, exist_ok=True)
    for name in sorted(os.listdir(SRC)):
        if name.endswith(".py"):
            a, b = os.path.join(SRC, name), os.path.join(DST, name)
            if not (os.path.exists(b) and filecmp.cmp(a, b, shallow=False)):
                shutil.copyfile(a, b)
                if verbose:
                    print("copied", name)
    return ref_root()


if __name__ == "__main__":
    print(make(verbose=True) or "reference tree not present", file=sys.stderr)
