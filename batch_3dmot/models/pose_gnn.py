"""Drop-in import path for the reference's batch_3dmot/models/pose_gnn.py."""
from batch3dmot_b200.pose_gnn import CausalMessagePassing, PoseGNN  # noqa: F401
