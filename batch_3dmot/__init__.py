"""Import-path shim: `batch_3dmot.models.{pose_gnn,clr_att_gnn}` resolve to the B200-native
layers in `batch3dmot_b200`, so train.py / predict.py style code keeps its imports."""
