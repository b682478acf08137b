"""Deterministic synthetic inputs: coordinates in [-1,1]^d, the latent grid of the
reference's data processor (src/datasets/data_processor.py:289-307: linspace + meshgrid('ij')), N(0,1)
features.  Pure functions of a seed, shared by bench.py, the tests and the golden generator."""
from __future__ import annotations

import torch


def latent_grid(size) -> torch.Tensor:
    axes = [torch.linspace(-1.0, 1.0, s, dtype=torch.float32) for s in size]
    return torch.stack(torch.meshgrid(*axes, indexing="ij"), dim=-1).reshape(-1, len(size)).contiguous()


def uniform_points(n: int, d: int, seed: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(n, d, generator=g, dtype=torch.float32) * 2.0 - 1.0).contiguous()


def clustered_points(n: int, d: int, seed: int, sigma: float = 0.15) -> torch.Tensor:
    """Airfoil-like density: half uniform, half N(0, sigma^2) around the origin, clipped to the box."""
    g = torch.Generator().manual_seed(seed)
    a = torch.rand(n // 2, d, generator=g, dtype=torch.float32) * 2.0 - 1.0
    b = (torch.randn(n - n // 2, d, generator=g, dtype=torch.float32) * sigma).clamp_(-1.0, 1.0)
    return torch.cat([a, b], dim=0).contiguous()


def normal(shape, seed: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g, dtype=torch.float32)
