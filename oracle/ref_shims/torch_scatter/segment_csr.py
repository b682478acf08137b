import torch


def segment_csr(src, indptr, out=None, reduce="sum"):
    """Reduce `src` along dim `indptr.dim()-1` over the CSR segments of `indptr` (broadcast over the
    leading dims).  Empty segments give 0; `mean` divides by max(count, 1); `max` returns values."""
    dim = indptr.dim() - 1
    ptr = indptr.reshape(-1, indptr.shape[-1])[0]
    n_seg = ptr.numel() - 1
    counts = ptr[1:] - ptr[:-1]
    seg = torch.repeat_interleave(torch.arange(n_seg, device=src.device), counts)
    shape = list(src.shape)
    shape[dim] = n_seg
    res = torch.zeros(shape, dtype=src.dtype, device=src.device)
    if reduce in ("sum", "mean"):
        res.index_add_(dim, seg, src)
        if reduce == "mean":
            view = [1] * src.dim()
            view[dim] = -1
            res = res / counts.clamp(min=1).to(src.dtype).view(view)
        return res
    if reduce == "max":
        view = [1] * src.dim()
        view[dim] = -1
        idx = seg.view(view).expand_as(src)
        return res.scatter_reduce(dim, idx, src, reduce="amax", include_self=False)
    raise ValueError(reduce)
