"""Dataset-side drop-ins (gaot_b200.datasets) against outputs of the reference's own GraphBuilder / DynamicPairDataset
(tests/golden/datasets.npz, written by oracle/make_golden.py from /root/reference/src/datasets)."""
import numpy as np
import pytest
import torch

from oracle import cases
from tests import abi_emulator
from tests.helpers import load_golden


def _check_graphs(g, tag, graphs):
    for b, per_scale in enumerate(graphs):
        assert isinstance(per_scale, list)
        for s, gr in enumerate(per_scale):
            assert set(k for k in gr if not k.startswith("_")) == {"neighbors_index", "neighbors_row_splits"}
            assert gr["neighbors_index"].dtype == torch.int64 and gr["neighbors_row_splits"].dtype == torch.int64
            assert np.array_equal(gr["neighbors_index"].cpu().numpy(), g[f"gb.{tag}.{b}.{s}.index"].astype(np.int64)), (tag, b, s)
            assert np.array_equal(gr["neighbors_row_splits"].cpu().numpy(), g[f"gb.{tag}.{b}.{s}.splits"].astype(np.int64)), (tag, b, s)


def test_graph_builder_matches_reference_host():
    """build_graphs_for_split through the ABI emulator: per-sample rescale, list[sample][scale] containers, bit-exact CSR."""
    from gaot_b200.datasets import GraphBuilder
    g = load_golden("datasets")
    x_data, lat, r, scales = cases.graph_builder_case()
    with abi_emulator.install():
        enc, dec = GraphBuilder("native").build_graphs_for_split(x_data, lat, r, scales)
        enc2, _ = GraphBuilder("native", samples_per_launch=3).build_graphs_for_split([x for x in x_data], lat, r, scales)
    assert len(enc) == len(dec) == x_data.shape[0] and all(len(e) == len(scales) for e in enc)
    _check_graphs(g, "enc", enc)
    _check_graphs(g, "dec", dec)
    _check_graphs(g, "enc", enc2)
    b = GraphBuilder("native")
    b.validate_graphs({"test": {"encoder": enc, "decoder": dec}, "train": None}, {"test": 4})
    with pytest.raises(AssertionError):
        b.validate_graphs({"test": {"encoder": enc, "decoder": dec}}, {"test": 5})
    with pytest.raises(ValueError):
        GraphBuilder("native").build_graphs_for_split(torch.zeros(2, 3, 5, 2, 1), lat, r, scales)


@pytest.mark.gpu
def test_graph_builder_matches_reference_gpu(tmp_path):
    from gaot_b200.datasets import CachedGraphBuilder
    g = load_golden("datasets")
    x_data, lat, r, scales = cases.graph_builder_case()
    b = CachedGraphBuilder("native", cache_dir=str(tmp_path))
    splits = {"test": {"x": x_data.cuda()}, "train": {"x": x_data[:2].cuda()}}
    graphs = b.build_all_graphs(splits, lat.cuda(), r, scales, dataset_name="fixture", build_train=True)
    assert graphs["test"]["encoder"][0][0]["neighbors_index"].is_cuda
    _check_graphs(g, "enc", graphs["test"]["encoder"])
    _check_graphs(g, "dec", graphs["test"]["decoder"])
    again = b.load_graphs("fixture", ["test", "train"])             # the reference's cache file names / format
    _check_graphs(g, "enc", again["test"]["encoder"])
    assert len(again["train"]["decoder"]) == 2


@pytest.mark.parametrize("vx", [False, True])
@pytest.mark.parametrize("mode", ["output", "residual", "time_der"])
def test_dynamic_pair_dataset_matches_reference(mode, vx):
    from gaot_b200.datasets import DynamicPairDataset
    g = load_golden("datasets")
    u, c, tv, x, stats = cases.pair_dataset_case()
    ds = DynamicPairDataset(u, c, tv, None, max_time_diff=6, time_step=2, stepper_mode=mode, stats=stats,
                            x_data=x if vx else None, is_variable_coords=vx)
    tag = f"pairs.{mode}.{'vx' if vx else 'fx'}"
    assert len(ds) == int(g[tag + ".len"])
    assert np.array_equal(ds.t_in_indices, g[tag + ".t_in"]) and np.array_equal(ds.t_out_indices, g[tag + ".t_out"])
    pick = g[tag + ".pick"].tolist()
    batch = ds.get_batch(pick)                                       # one batched gather
    assert torch.allclose(batch[0], torch.from_numpy(g[tag + ".input"]), rtol=0, atol=1e-6)
    assert torch.allclose(batch[1], torch.from_numpy(g[tag + ".target"]), rtol=1e-6, atol=1e-6)
    if vx:
        assert torch.equal(batch[2], torch.from_numpy(g[tag + ".x"]))
    item = ds[pick[1]]                                               # the per-item contract of the reference
    assert len(item) == (3 if vx else 2) and torch.allclose(item[0], batch[0][1]) and torch.allclose(item[1], batch[1][1])


def test_dynamic_pair_dataset_without_stats():
    from gaot_b200.datasets import DynamicPairDataset
    g = load_golden("datasets")
    u, _, tv, _, _ = cases.pair_dataset_case()
    ds = DynamicPairDataset(u, None, tv, None, max_time_diff=14, time_step=1, stepper_mode="residual", stats=None, use_time_norm=False)
    assert len(ds) == int(g["pairs.nostats.len"])
    batch = ds.get_batch(list(range(0, len(ds), 7)))
    assert torch.allclose(batch[0], torch.from_numpy(g["pairs.nostats.input"]), atol=1e-6)
    assert torch.allclose(batch[1], torch.from_numpy(g["pairs.nostats.target"]), atol=1e-6)
    with pytest.raises(ValueError):
        DynamicPairDataset(u, None, tv, None, stepper_mode="euler")
