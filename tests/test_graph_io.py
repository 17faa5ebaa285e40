"""Graph file I/O: batch3dmot_b200.graph_io.load_window_graph against the literal restatement of
GraphDataset.__getitem__ (oracle/graph_io.py) on file sets written in the reference's layout. Everything bit-exact."""
import os

import pytest
import torch

from oracle import graph_io as R
from batch3dmot_b200 import graph_io, synth


def write_window(tmp, seed, T=5, npf=12):
    scene = synth.add_labels(synth.add_modalities(synth.scene_graph(seed=seed, T=T, nodes_per_frame=npf, k=8, rel_time_mod=5), seed), seed)
    prefix = os.path.join(str(tmp), f"scene{seed}_len5_0")
    meta = {n: {"category_name": synth.CATEGORIES[int(c) - 1], "global_node_id": 1000 + 3 * n, "time": int(t)}
            for n, (c, t) in enumerate(zip(scene.node_classes.tolist(), scene.node_timestamps.tolist()))}
    boxes = torch.randn(scene.num_nodes, 10)
    graph_io.save_window_graph(prefix, scene, meta, boxes)
    return prefix, scene


def same(a, b):
    return a.dtype == b.dtype and a.shape == b.shape and torch.equal(a, b)


@pytest.mark.parametrize("seed,inference", [(1, False), (2, True), (3, True)])
def test_loader_matches_reference_getitem(tmp_path, seed, inference):
    prefix, scene = write_window(tmp_path, seed)
    ref = R.getitem(prefix, inference=inference)
    got = graph_io.load_window_graph(prefix, inference=inference)
    keys = ["pose_feats", "img_feats", "lidar_feats", "radar_feats", "edge_index", "edge_attr", "y", "node_timestamps",
            "edge_weights", "edge_classes", "node_classes"] + (["global_edge_index", "global_node_timestamps", "boxes"] if inference else [])
    for k in keys:
        assert same(getattr(got, k), getattr(ref, k)), k
    assert got.num_nodes == ref.num_nodes == scene.num_nodes
    assert got.edge_index.shape[0] == 2 and got.edge_attr.dtype == torch.float64     # `.float()` happens in the model


def test_no_weighting_and_empty_graph(tmp_path):
    prefix, _ = write_window(tmp_path, 4)
    ref, got = R.getitem(prefix, edge_weighting=False), graph_io.load_window_graph(prefix, edge_weighting=False)
    assert same(got.edge_weights, ref.edge_weights) and got.edge_classes is None
    prefix, scene = write_window(tmp_path, 5, T=1)           # a single frame: no edges
    assert scene.edge_index.size(1) == 0
    got = graph_io.load_window_graph(prefix, inference=True)
    assert got.edge_index.shape == (2, 0) and got.edge_weights.numel() == 0 and got.global_edge_index.shape == (2, 0)


def test_mixed_category_edges_are_rejected_like_the_reference(tmp_path):
    prefix, scene = write_window(tmp_path, 6)
    meta = {n: {"category_name": "car" if n % 2 else "bus", "global_node_id": n} for n in range(scene.num_nodes)}
    graph_io.save_window_graph(prefix, scene, meta)
    with pytest.raises(AttributeError):
        R.getitem(prefix)
    with pytest.raises(NotImplementedError):
        graph_io.load_window_graph(prefix)


def test_load_batch_collates_windows(tmp_path):
    p1, s1 = write_window(tmp_path, 7)
    p2, s2 = write_window(tmp_path, 8, npf=9)
    b = graph_io.load_batch([p1, p2], pin=False)
    assert b.num_nodes == s1.num_nodes + s2.num_nodes
    E1 = s1.edge_index.size(1)
    assert torch.equal(b.edge_index[:, E1:], s2.edge_index + s1.num_nodes)
    assert torch.equal(b.batch, torch.cat([torch.zeros(s1.num_nodes), torch.ones(s2.num_nodes)]).long())
    assert b.edge_weights.shape == (b.edge_index.size(1),) and b.pose_feats.shape == (b.num_nodes, 19)


def test_window_batch_loader_order_and_content(tmp_path):
    prefixes = [write_window(tmp_path, 40 + i, npf=6 + i)[0] for i in range(5)]
    ld = graph_io.WindowBatchLoader(prefixes, batch_size=2, workers=3, prefetch=2, pin=False)
    assert len(ld) == 3
    got = list(ld)
    assert len(got) == 3
    for b, grp in zip(got, ld.batches()):
        ref = graph_io.load_batch(grp, pin=False)
        for k in ("pose_feats", "edge_index", "edge_attr", "y", "edge_weights", "batch"):
            assert torch.equal(getattr(b, k), getattr(ref, k)), k
    assert int(got[-1].batch.max()) == 0                                # the last batch holds the single left-over window
    assert len(graph_io.WindowBatchLoader(prefixes, batch_size=2, drop_last=True)) == 2
    sh = graph_io.WindowBatchLoader(prefixes, batch_size=2, shuffle=True, seed=3, pin=False)
    e0, _ = sh.batches(), list(sh)
    e1 = sh.batches()
    assert sorted(sum(e0, [])) == sorted(prefixes) and e0 != e1          # a permutation, reshuffled every epoch
    sh2 = graph_io.WindowBatchLoader(prefixes, batch_size=2, shuffle=True, seed=3, pin=False)
    assert sh2.batches() == e0                                           # deterministic for a given seed


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree only exists in the build container")
@pytest.mark.parametrize("inference", [False, True])
def test_live_unmodified_reference_getitem(tmp_path, inference):
    """The UNMODIFIED GraphDataset.__getitem__ (graph_data.py:152-256, imported from /root/reference under the shim)
    on the same files: pins both the restatement and the loader."""
    from types import SimpleNamespace
    from oracle import pyg_shim
    GD = pyg_shim.load_reference_graph_dataset()
    prefix, _ = write_window(tmp_path, 9)
    ds = object.__new__(GD)                      # __init__ only builds the file list from nuScenes scene records
    ds.batches, ds.inference, ds.edge_weighting = [prefix], inference, True
    ds.params = SimpleNamespace(classes=SimpleNamespace(nuscenes_tracking_eval=R.CLASS_DICT),
                                main=SimpleNamespace(class_dict="nuscenes_tracking_eval"))
    ds.rel_freq_train = R.REL_FREQ_TRAIN
    item = ds[0]
    data = item[0] if inference else item
    ours, rest = graph_io.load_window_graph(prefix, inference=inference), R.getitem(prefix, inference=inference)
    keys = ["pose_feats", "img_feats", "lidar_feats", "radar_feats", "edge_index", "edge_attr", "y", "node_timestamps",
            "edge_weights", "edge_classes", "node_classes"] + (["global_edge_index", "global_node_timestamps", "boxes"] if inference else [])
    for k in keys:
        assert same(getattr(data, k), getattr(ours, k)) and same(getattr(data, k), getattr(rest, k)), k
    assert data.num_nodes == ours.num_nodes


@pytest.mark.gpu
def test_window_loader_feeds_the_gpu_from_pinned_batches_with_overlapped_copies(tmp_path):
    """train.py:85-97 on the box: background threads read + collate window files into PINNED batches; the consumer
    copies batch i+1 on a copy stream while the model runs batch i. Results equal the synchronous path bit for bit,
    and the copy stream really ran concurrently with compute (its copies do not serialise behind the forward)."""
    from types import SimpleNamespace
    from batch3dmot_b200 import ops
    from batch3dmot_b200.pose_gnn import PoseGNN
    dev = torch.device("cuda")
    prefixes = [write_window(tmp_path, 60 + i, npf=20 + i)[0] for i in range(12)]
    torch.manual_seed(5621)
    model = PoseGNN().to(dev).eval()
    keys = ("pose_feats", "edge_index", "edge_attr", "node_timestamps", "batch", "y", "edge_weights")

    def forward(b):
        with torch.no_grad():
            return model(b)[0]

    # synchronous reference
    ref = []
    for grp in graph_io.WindowBatchLoader(prefixes, batch_size=2, pin=False).batches():
        b = graph_io.load_batch(grp, pin=False)
        ref.append(forward(SimpleNamespace(**{k: getattr(b, k).to(dev) for k in keys})).cpu())
    # pipelined: pinned batches, non_blocking H2D on a copy stream, event hand-over to the compute stream
    copy = torch.cuda.Stream(device=dev)
    outs, staged = [], None
    loader = graph_io.WindowBatchLoader(prefixes, batch_size=2, workers=4, prefetch=3, pin=True)

    def stage(b):
        assert all(getattr(b, k).is_pinned() for k in keys)
        with torch.cuda.stream(copy):
            d = SimpleNamespace(**{k: getattr(b, k).to(dev, non_blocking=True) for k in keys})
            ev = torch.cuda.Event()
            ev.record(copy)
        return d, ev, b                                   # keep the pinned source alive until the copy is consumed

    for b in loader:
        nxt = stage(b)                                    # H2D of this batch is in flight while the previous one computes
        if staged is not None:
            d, ev, _ = staged
            torch.cuda.current_stream().wait_event(ev)
            outs.append(forward(d))
        staged = nxt
    d, ev, _ = staged
    torch.cuda.current_stream().wait_event(ev)
    outs.append(forward(d))
    torch.cuda.synchronize()
    assert len(outs) == len(ref) == 6
    for a, r in zip(outs, ref):
        assert torch.equal(a.cpu(), r)
