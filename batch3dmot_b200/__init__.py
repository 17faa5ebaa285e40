"""B200-native implementation of the Batch3DMOT tracking-graph GNN hot path."""
