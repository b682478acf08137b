"""Train-time edge sampling on CSR graphs (reference src/model/layers/utils/edge_drop.py:8-105).

Identity unless training with a sampling strategy (default `None`, magno.py:51).  The sampling itself
is host-orchestrated torch RNG work on the integer CSR arrays (as in the reference); it is not on the
benchmarked path and draws from torch's generator, so masks differ from the reference's only through
RNG consumption order."""
from __future__ import annotations

import torch


def apply_edge_drop_csr(neighbors, sampling_strategy=None, max_neighbors=None, sample_ratio=None, training=True):
    if not training or sampling_strategy is None:
        return neighbors
    idx, rs = neighbors["neighbors_index"], neighbors["neighbors_row_splits"]
    deg = rs[1:] - rs[:-1]
    E = idx.numel()
    if sampling_strategy == "ratio":
        if sample_ratio is None or not 0 < sample_ratio <= 1:
            raise ValueError("sample_ratio must be in (0, 1] when using 'ratio' sampling")
        keep = torch.rand(E, device=idx.device) < sample_ratio
    elif sampling_strategy == "max_neighbors":
        if max_neighbors is None or max_neighbors <= 0:
            raise ValueError("max_neighbors must be > 0 when using 'max_neighbors' sampling")
        # random rank of every edge inside its row; keep the max_neighbors smallest ranks
        row = torch.repeat_interleave(torch.arange(deg.numel(), device=idx.device), deg)
        key = torch.rand(E, device=idx.device) + row.to(torch.float32) * 2.0
        order = torch.argsort(key)
        rank = torch.empty_like(order)
        rank[order] = torch.arange(E, device=idx.device)
        keep = (rank - rs[:-1][row]) < max_neighbors
    else:
        raise ValueError(f"Unknown sampling_strategy: {sampling_strategy}")
    row = torch.repeat_interleave(torch.arange(deg.numel(), device=idx.device), deg)
    new_deg = torch.zeros_like(deg).index_add_(0, row[keep], torch.ones_like(row[keep]))
    new_rs = torch.zeros_like(rs)
    new_rs[1:] = torch.cumsum(new_deg, 0)
    return {"neighbors_index": idx[keep].contiguous(), "neighbors_row_splits": new_rs}
