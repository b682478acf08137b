"""CPU stand-in for the `torch_scatter` wheel (absent in this image), following its documented
semantics: empty segments/bins reduce to 0, `mean` divides by max(count, 1), `scatter_max` returns
(values, argmax).  Used only to run the unmodified reference as a golden-vector generator."""
import torch
from .segment_csr import segment_csr  # noqa: F401  (rebinds the submodule name to the function, like the wheel)


def _expand_index(index, src, dim):
    if index.dim() == 1 and src.dim() > 1:
        shape = [1] * src.dim()
        shape[dim] = -1
        index = index.view(shape).expand_as(src)
    return index


def scatter_sum(src, index, dim=0, out=None, dim_size=None):
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    shape = list(src.shape)
    shape[dim] = dim_size
    res = torch.zeros(shape, dtype=src.dtype, device=src.device)
    return res.scatter_add_(dim, _expand_index(index, src, dim), src)


def scatter_mean(src, index, dim=0, out=None, dim_size=None):
    total = scatter_sum(src, index, dim, None, dim_size)
    ones = torch.ones(index.shape[0], dtype=src.dtype, device=src.device)
    cnt = scatter_sum(ones, index, 0, None, total.shape[dim]).clamp_(min=1)
    shape = [1] * total.dim()
    shape[dim] = -1
    return total / cnt.view(shape)


def scatter_max(src, index, dim=0, out=None, dim_size=None):
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    shape = list(src.shape)
    shape[dim] = dim_size
    idx = _expand_index(index, src, dim)
    res = torch.zeros(shape, dtype=src.dtype, device=src.device)
    res = res.scatter_reduce(dim, idx, src, reduce="amax", include_self=False)
    arg = torch.full(shape, src.shape[dim], dtype=torch.long, device=src.device)
    return res, arg
