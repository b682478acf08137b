"""Train-time edge sampling on CSR graphs (reference src/model/layers/utils/edge_drop.py:8-105).

Identity unless training with a sampling strategy (default `None`, magno.py:51).  Same contract as the
reference, including its permissive early returns: no edges, `sample_ratio` None or >= 1, `max_neighbors`
None, no row above the cap, and an unknown strategy all hand the input graph back unchanged
(edge_drop.py:41-42, 55-56, 71-72, 79-80, 98-99); a corrupt CSR raises ValueError (edge_drop.py:58-60, 74-76).

The sampling is integer / RNG work on the CSR arrays on whatever device they live on (no host loop: the
reference's per-row `randperm` loop, edge_drop.py:84-92, becomes one two-key sort).  Masks are drawn from
torch's generator, so they differ from the reference's through RNG consumption order only; the
distribution is the same (uniform subset of exactly `max_neighbors` edges per over-full row; i.i.d.
Bernoulli(sample_ratio) per edge) - tests/test_host_logic.py checks it statistically."""
from __future__ import annotations

import torch


def apply_edge_drop_csr(neighbors, sampling_strategy=None, max_neighbors=None, sample_ratio=None, training=True):
    if not training or sampling_strategy is None:
        return neighbors
    idx, rs = neighbors["neighbors_index"], neighbors["neighbors_row_splits"]
    E = idx.numel()
    if E == 0:
        return neighbors
    m = rs.shape[0] - 1
    dev = idx.device
    if sampling_strategy == "ratio":
        if sample_ratio is None or sample_ratio >= 1.0:
            return neighbors
        deg = rs[1:] - rs[:-1]
        if bool((deg < 0).any()) or int(deg.sum()) != E:
            raise ValueError("Invalid CSR structure before ratio sampling.")
        keep = torch.rand(E, device=dev) < sample_ratio
        row = torch.repeat_interleave(torch.arange(m, device=dev), deg)
        new_deg = torch.bincount(row[keep], minlength=m)
    elif sampling_strategy == "max_neighbors":
        if max_neighbors is None:
            return neighbors
        deg = rs[1:] - rs[:-1]
        if bool((deg < 0).any()) or int(deg.sum()) != E:
            raise ValueError("Invalid CSR structure before max_neighbors sampling.")
        if not bool((deg > max_neighbors).any()):
            return neighbors
        # a uniformly random rank of every edge inside its row: sort by an i.i.d. key, then stably by row, so the
        # randomness never shares a floating-point word with the row number (exact for any number of rows)
        row = torch.repeat_interleave(torch.arange(m, device=dev), deg)
        by_key = torch.argsort(torch.rand(E, device=dev, dtype=torch.float64))
        order = by_key[torch.sort(row[by_key], stable=True).indices]      # edges grouped by row, random order inside a row
        rank = torch.empty(E, dtype=torch.int64, device=dev)
        rank[order] = torch.arange(E, device=dev) - rs[:-1].to(torch.int64)[row[order]]
        keep = rank < max_neighbors
        new_deg = torch.clamp(deg, max=max_neighbors)
    else:
        return neighbors
    new_rs = torch.zeros(m + 1, dtype=rs.dtype, device=dev)
    new_rs[1:] = torch.cumsum(new_deg, 0)
    return {"neighbors_index": idx[keep], "neighbors_row_splits": new_rs}
