"""Drop-in import path for the reference's batch_3dmot/models/clr_att_gnn.py."""
from batch3dmot_b200.clr_att_gnn import GNN, CausalMessagePassing  # noqa: F401
