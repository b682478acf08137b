"""Drop-in for reference src/model/layers/utils/neighbor_search.py (`NeighborSearch`, :16-106).

Every `method` resolves to the same sm_100a kernel pair (count + fill); the method name only selects
the *semantics* the reference back-end of that name has:
  native / chunked / grid / auto / open3d : distance <= r, all neighbours      (neighbor_search.py:133-136)
  torch_cluster                           : distance <  r, first 32 neighbours (neighbor_search.py:163-165)
Rows are always in ascending neighbour index (the reference's `grid` back-end returns the same sets in
cell order; its docstring at :84-86 allows a permutation).
"""
from __future__ import annotations

import torch
from torch import nn

from ....ops import radius_csr

_METHODS = ("open3d", "torch_cluster", "grid", "chunked", "native", "auto")


class NeighborSearch(nn.Module):
    def __init__(self, method: str = "auto", return_norm: bool = False):
        super().__init__()
        if method not in _METHODS:
            raise ValueError(f"Unknown neighbor search method: {method}")
        self.method = method
        self.return_norm = return_norm

    def forward(self, data: torch.Tensor, queries: torch.Tensor, radius: float):
        if self.method == "torch_cluster":
            return radius_csr(data, queries, radius, inclusive=False, max_num_neighbors=32)
        return radius_csr(data, queries, radius, inclusive=True, max_num_neighbors=None)
