"""Drop-in for reference src/datasets/graph_builder.py (`GraphBuilder`, :13-173; `CachedGraphBuilder`, :176-287).

Same constructor, methods, arguments and return containers: `build_graphs_for_split` returns
(encoder_graphs, decoder_graphs), each `list[sample][scale]` of {"neighbors_index", "neighbors_row_splits"} dicts.
B200-native difference: the reference loops over samples in Python and runs one neighbour search per sample and scale
(graph_builder.py:46-78); here the per-sample min-max rescale to [-1, 1] (graph_builder.py:60, utils/scaling.py:10-36) is
one batched reduction and every scale needs ONE batched count + fill launch pair for all samples of a chunk
(ops.radius_csr with a leading sample dimension); the graphs stay on the device the coordinates live on."""
from __future__ import annotations

import os
from typing import List, Optional, Tuple

import torch

from ..model.layers.utils.neighbor_search import NeighborSearch
from .. import ops


def rescale(data: torch.Tensor, target_range: Tuple[float, float] = (-1, 1)) -> torch.Tensor:
    """utils/scaling.py:10-36 for [n, d] or batched [S, n, d] input (min / max over the point axis), same arithmetic order:
    (data - min) / range * (hi - lo) + lo with a zero range replaced by 1."""
    dim = data.ndim - 2
    data_min = torch.min(data, dim=dim, keepdim=True)[0]
    data_max = torch.max(data, dim=dim, keepdim=True)[0]
    data_range = data_max - data_min
    data_range = torch.where(data_range == 0, torch.ones_like(data_range), data_range)
    normalized = (data - data_min) / data_range
    target_min, target_max = target_range
    return normalized * (target_max - target_min) + target_min


class GraphBuilder:
    def __init__(self, neighbor_search_method: str = "auto", samples_per_launch: int = 256):
        self.nb_search = NeighborSearch(neighbor_search_method)
        self.samples_per_launch = samples_per_launch

    def _search_kwargs(self):
        if self.nb_search.method == "torch_cluster":
            return dict(inclusive=False, max_num_neighbors=32)
        return dict(inclusive=True, max_num_neighbors=None)

    def build_graphs_for_split(self, x_data: torch.Tensor, latent_queries: torch.Tensor, gno_radius: float,
                               scales: List[float]) -> Tuple[List, List]:
        """x_data [n_samples, n_nodes, d] or [n_samples, 1, n_nodes, d] -> (encoder_graphs, decoder_graphs)."""
        if torch.is_tensor(x_data):
            if x_data.ndim == 4 and x_data.shape[1] == 1:
                x = x_data[:, 0]
            elif x_data.ndim == 3:
                x = x_data
            else:
                raise ValueError(f"Unexpected coordinate shape: {tuple(x_data.shape[1:])}")
        else:                                               # a list of per-sample tensors of equal size
            xs = []
            for x_sample in x_data:
                if x_sample.dim() == 3 and x_sample.shape[0] == 1:
                    xs.append(x_sample[0])
                elif x_sample.dim() == 2:
                    xs.append(x_sample)
                else:
                    raise ValueError(f"Unexpected coordinate shape: {x_sample.shape}")
            x = torch.stack(xs)
        dev = latent_queries.device
        lat = latent_queries.to(torch.float32).contiguous()
        kw = self._search_kwargs()
        encoder_graphs: List[list] = [[] for _ in range(x.shape[0])]
        decoder_graphs: List[list] = [[] for _ in range(x.shape[0])]
        with torch.no_grad():
            for s0 in range(0, x.shape[0], self.samples_per_launch):
                xc = rescale(x[s0:s0 + self.samples_per_launch].to(dev, torch.float32), (-1, 1)).contiguous()
                for scale in scales:
                    r = gno_radius * scale
                    enc = ops.radius_csr(xc, lat, r, **kw)             # physical -> latent, one launch pair per chunk
                    dec = ops.radius_csr(lat, xc, r, **kw)             # latent -> physical
                    for i in range(xc.shape[0]):
                        encoder_graphs[s0 + i].append(enc[i])
                        decoder_graphs[s0 + i].append(dec[i])
        return encoder_graphs, decoder_graphs

    def build_all_graphs(self, data_splits: dict, latent_queries: torch.Tensor, gno_radius: float, scales: List[float],
                         build_train: bool = True) -> dict:
        all_graphs = {}
        if "test" in data_splits:
            enc, dec = self.build_graphs_for_split(data_splits["test"]["x"], latent_queries, gno_radius, scales)
            all_graphs["test"] = {"encoder": enc, "decoder": dec}
        if build_train:
            for split in ("train", "val"):
                if split in data_splits:
                    enc, dec = self.build_graphs_for_split(data_splits[split]["x"], latent_queries, gno_radius, scales)
                    all_graphs[split] = {"encoder": enc, "decoder": dec}
        else:
            all_graphs["train"] = None
            all_graphs["val"] = None
        return all_graphs

    def validate_graphs(self, graphs: dict, expected_samples: dict):
        for split_name, split_graphs in graphs.items():
            if split_graphs is None:
                continue
            enc, dec = split_graphs["encoder"], split_graphs["decoder"]
            expected = expected_samples.get(split_name, 0)
            assert len(enc) == expected, f"Encoder graphs for {split_name}: expected {expected}, got {len(enc)}"
            assert len(dec) == expected, f"Decoder graphs for {split_name}: expected {expected}, got {len(dec)}"
            for i, (e, d) in enumerate(zip(enc, dec)):
                assert isinstance(e, list), f"Encoder sample {i} should be list of scales"
                assert isinstance(d, list), f"Decoder sample {i} should be list of scales"
                assert len(e) == len(d), f"Encoder and decoder should have same number of scales for sample {i}"


class CachedGraphBuilder(GraphBuilder):
    """graph_builder.py:176-287: same cache file names and torch.save format (graphs are moved to the CPU for saving)."""

    def __init__(self, neighbor_search_method: str = "auto", cache_dir: Optional[str] = None):
        super().__init__(neighbor_search_method)
        self.cache_dir = cache_dir

    def _get_cache_path(self, dataset_name: str, split_name: str, graph_type: str) -> str:
        if self.cache_dir is None:
            raise ValueError("Cache directory not specified")
        os.makedirs(self.cache_dir, exist_ok=True)
        return os.path.join(self.cache_dir, f"{dataset_name}_{split_name}_{graph_type}_graphs.pt")

    @staticmethod
    def _cpu(graphs):
        return [[{k: v.cpu() for k, v in g.items() if torch.is_tensor(v)} for g in per_scale] for per_scale in graphs]

    def save_graphs(self, graphs: dict, dataset_name: str):
        if self.cache_dir is None:
            return
        for split_name, split_graphs in graphs.items():
            if split_graphs is None:
                continue
            torch.save(self._cpu(split_graphs["encoder"]), self._get_cache_path(dataset_name, split_name, "encoder"))
            torch.save(self._cpu(split_graphs["decoder"]), self._get_cache_path(dataset_name, split_name, "decoder"))

    def load_graphs(self, dataset_name: str, splits: List[str]) -> Optional[dict]:
        if self.cache_dir is None:
            return None
        try:
            out = {}
            for split_name in splits:
                ep = self._get_cache_path(dataset_name, split_name, "encoder")
                dp = self._get_cache_path(dataset_name, split_name, "decoder")
                if not (os.path.exists(ep) and os.path.exists(dp)):
                    return None
                out[split_name] = {"encoder": torch.load(ep), "decoder": torch.load(dp)}
            return out
        except Exception:
            return None

    def build_all_graphs(self, data_splits: dict, latent_queries: torch.Tensor, gno_radius: float, scales: List[float],
                         dataset_name: str = "dataset", build_train: bool = True, use_cache: bool = True) -> dict:
        if use_cache and self.cache_dir is not None:
            splits = ["test"] + (["train", "val"] if build_train else [])
            cached = self.load_graphs(dataset_name, splits)
            if cached is not None:
                return cached
        graphs = super().build_all_graphs(data_splits, latent_queries, gno_radius, scales, build_train)
        if use_cache:
            self.save_graphs(graphs, dataset_name)
        return graphs
