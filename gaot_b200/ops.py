"""Host-side operators over the C ABI (include/gaot_b200.h): thin torch.autograd.Function wrappers
whose forward AND backward are calls into libgaot_b200.so.  torch is used for device memory, the
stream and the autograd tape only - there is no eager/CPU fallback behind any of these.

Shapes are row-major contiguous fp32; CSR arrays are int64 as in the reference container
({"neighbors_index", "neighbors_row_splits"}, reference src/model/layers/utils/neighbor_search.py:80-90).
"""
from __future__ import annotations

import os

from typing import Dict, List, Optional

import torch

from . import _lib
from ._lib import call, ptr

ACT_NONE, ACT_GELU, ACT_RELU = 0, 1, 2

# Numerics mode of the processor's dense layers and attention.
#   "fp32": the parity path (<= 1e-5 vs the reference, which computes in fp32) ON THE TENSOR CORES: every operand is split
#           into fp16 hi + lo planes (~22 bits) and each product runs as three tcgen05 terms with fp32 accumulation
#           (gaot_gemm_split / gaot_attention_*_split); shapes the split kernels do not take fall back to the FFMA kernels.
#   "ffma": fp32 FFMA kernels everywhere (the round-1 parity path; kept as the cross-check of the split kernels).
#   "bf16": tcgen05 kernels with bf16 operands and fp32 accumulation, the performance path (tolerances in tests/test_gpu_tc.py).
# MAGNO's gather / reduce kernels stay fp32 in every mode.
_PRECISION = "fp32"


def set_precision(mode: str):
    global _PRECISION
    if mode not in ("fp32", "ffma", "bf16"):
        raise ValueError("precision must be 'fp32', 'ffma' or 'bf16'")
    _PRECISION = mode


def get_precision() -> str:
    return _PRECISION


def _c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


def _need_cuda(t: torch.Tensor):
    if not t.is_cuda:
        raise _lib.GaotLibraryError("gaot_b200 runs on CUDA tensors only - there is no CPU fallback")


# ------------------------------------------------------------------------------------------------
# graphs
# ------------------------------------------------------------------------------------------------
GRID_SEARCH_MIN_PAIRS = 1 << 22      # n * m above which the cell-binned search replaces the O(n m) sweep (None cap only)


def _grid_geometry(data, radius):
    """Uniform grid over the bounding box of the data points: pitch >= 1.001 r, about two points per cell, at most 2^22
    cells.  One small device -> host read (the box); the graph build reads E back anyway."""
    d = data.shape[-1]
    flat = data.reshape(-1, d)
    box = torch.stack([flat.amin(0), flat.amax(0)]).cpu()
    lo, hi = box[0].tolist(), box[1].tolist()
    n = data.shape[-2]
    ext = [max(h_ - l_, 0.0) for l_, h_ in zip(lo, hi)]
    vol = 1.0
    for e in ext:
        vol *= max(e, 1e-12)
    h = max(1.001 * float(radius), (2.0 * vol / max(n, 1)) ** (1.0 / d))
    while True:
        g = [int(e / h) + 1 for e in ext]
        total = 1
        for v in g:
            total *= v
        if total <= (1 << 22):
            break
        h *= 1.26
    return lo + [0.0] * (3 - d), h, g + [1] * (3 - d)


def radius_csr(data: torch.Tensor, queries: torch.Tensor, radius: float, inclusive: bool = True,
               max_num_neighbors: Optional[int] = None, method: str = "auto") -> Dict[str, torch.Tensor]:
    """Fixed-radius graph in the reference's CSR container (neighbor_search.py:65-106).
    data [n,d] (or [S,n,d]), queries [m,d] (or [S,m,d]); a batched side yields a list of S dicts.
    method: 'sweep' (O(n m) warp sweep), 'grid' (cell-binned), 'auto' (grid for large problems without a neighbour cap);
    both produce the same bit-exact graph."""
    _need_cuda(data)
    data, queries = _c(data), _c(queries)
    db, qb = data.ndim == 3, queries.ndim == 3
    S = data.shape[0] if db else (queries.shape[0] if qb else 1)
    n, m, d = data.shape[-2], queries.shape[-2], data.shape[-1]
    cap = 0 if max_num_neighbors is None else int(max_num_neighbors)
    if method not in ("auto", "sweep", "grid"):
        raise ValueError(f"unknown radius search method {method!r}")
    if method == "grid" and cap:
        raise ValueError("the cell-binned search returns all neighbours (max_num_neighbors must be None)")
    use_grid = n > 0 and m > 0 and (method == "grid" or (method == "auto" and cap == 0 and n * m >= GRID_SEARCH_MIN_PAIRS))
    splits = torch.empty((S, m + 1), dtype=torch.int64, device=data.device)
    if use_grid:
        lo, h, g = _grid_geometry(data, radius)
        ncell = g[0] * g[1] * g[2]
        Sd = S if db else 1
        i32 = dict(dtype=torch.int32, device=data.device)
        cell_start, srt = torch.empty((Sd, ncell + 1), **i32), torch.empty((Sd, n), **i32)
        cell_ws, cursor_ws = torch.empty((Sd, n), **i32), torch.empty((Sd, ncell), **i32)     # distinct live workspaces
        call("gaot_grid_bin", ptr(data), n, d, Sd, lo[0], lo[1], lo[2], float(h), g[0], g[1], g[2], ptr(cell_start), ptr(srt),
             ptr(cell_ws), ptr(cursor_ws))
        geom = (lo[0], lo[1], lo[2], float(h), g[0], g[1], g[2], ptr(cell_start), ptr(srt))
        call("gaot_radius_count_grid", ptr(data), n, ptr(queries), m, d, float(radius), int(inclusive), S, int(db), int(qb), *geom,
             ptr(splits))
    else:
        call("gaot_radius_count", ptr(data), n, ptr(queries), m, d, float(radius), int(inclusive), cap, S, int(db), int(qb),
             ptr(splits))
    counts = splits[:, m].clone()
    offs = torch.zeros(S, dtype=torch.int64, device=data.device)
    if S > 1:
        offs[1:] = torch.cumsum(counts, 0)[:-1]
    counts_h = counts.cpu()                       # E is data dependent: one host read per build
    total = int(counts_h.sum())
    index = torch.empty(max(total, 1), dtype=torch.int64, device=data.device)[:total]
    if total > 0:
        if use_grid:
            call("gaot_radius_fill_grid", ptr(data), n, ptr(queries), m, d, float(radius), int(inclusive), S, int(db), int(qb), *geom,
                 ptr(splits), ptr(offs), ptr(index))
        else:
            call("gaot_radius_fill", ptr(data), n, ptr(queries), m, d, float(radius), int(inclusive), cap, S, int(db), int(qb),
                 ptr(splits), ptr(offs), ptr(index))
    out, o = [], 0
    for s in range(S):
        e = int(counts_h[s])
        out.append({"neighbors_index": index[o:o + e], "neighbors_row_splits": splits[s]})
        o += e
    return out if (db or qb) else out[0]


class GraphPlan:
    """Device-side companions of one CSR graph that the kernels need: the row of every edge and the
    source-sorted (transposed) view used by the backward of the gather."""

    def __init__(self, index: torch.Tensor, splits: torch.Tensor, n_src: int):
        _need_cuda(index)
        self.index = index.to(torch.int64).contiguous()
        self.splits = splits.to(torch.int64).contiguous()
        self.n_src = int(n_src)
        self.m = int(self.splits.numel() - 1)
        self.E = int(self.index.numel())
        dev = self.index.device
        self.edge_row = torch.empty(max(self.E, 1), dtype=torch.int64, device=dev)[:self.E]
        call("gaot_csr_edge_rows", ptr(self.splits), self.m, ptr(self.edge_row))
        self._t = None
        self._geo: Dict = {}
        self._slices: Dict = {}

    def row_slice(self, part: int, parts: int) -> "GraphPlan":
        """Sub-graph of the destination rows [part * m / parts, (part + 1) * m / parts) as a plan of its own (sources keep
        their numbering); needs m % parts == 0.  Cached (one host read of two row offsets on first use)."""
        key = (part, parts)
        hit = self._slices.get(key)
        if hit is None:
            per = self.m // parts
            lo, hi = part * per, (part + 1) * per
            a, b = (int(v) for v in self.splits[[lo, hi]].cpu())
            hit = GraphPlan(self.index[a:b], self.splits[lo:hi + 1] - a, self.n_src)
            hit.row_lo, hit.row_hi, hit.edge_lo, hit.edge_hi = lo, hi, a, b
            self._slices[key] = hit
        return hit

    def transposed(self):
        if self._t is None:
            dev = self.index.device
            ts = torch.empty(self.n_src + 1, dtype=torch.int64, device=dev)
            tr = torch.empty(max(self.E, 1), dtype=torch.int64, device=dev)[:self.E]
            tp = torch.empty(max(self.E, 1), dtype=torch.int64, device=dev)[:self.E]
            ws = torch.empty(max(self.n_src, 1), dtype=torch.int64, device=dev)
            call("gaot_csr_transpose", ptr(self.index), ptr(self.splits), self.m, self.n_src, self.E, ptr(ts), ptr(tr),
                 ptr(tp), ptr(ws))
            self._t = (ts, tr, tp)
        return self._t


def graph_plan(nbrs, n_src: int) -> GraphPlan:
    """Plan cached inside the graph dict (key '_gaot_plan') so that it lives exactly as long as the graph."""
    if isinstance(nbrs, GraphPlan):
        return nbrs
    plan = nbrs.get("_gaot_plan") if isinstance(nbrs, dict) else None
    if plan is None or plan.n_src != n_src or plan.index.data_ptr() != nbrs["neighbors_index"].data_ptr():
        plan = GraphPlan(nbrs["neighbors_index"], nbrs["neighbors_row_splits"], n_src)
        if isinstance(nbrs, dict):
            nbrs["_gaot_plan"] = plan
    return plan


def merge_graphs(graphs: List[dict], n_src_each: int) -> GraphPlan:
    """Block-diagonal merge of per-sample CSR graphs (vx mode): sample b's sources are offset by
    b*n_src_each and its rows appended after sample b-1's rows.  One launch then serves the batch.
    A fixed number of device ops whatever the batch size, and no device -> host read: the per-sample edge counts are
    tensor shapes, i.e. already on the host."""
    B = len(graphs)
    dev = graphs[0]["neighbors_index"].device
    Es = [int(g["neighbors_index"].numel()) for g in graphs]
    E = sum(Es)
    idx = torch.cat([g["neighbors_index"] for g in graphs]).to(torch.int64)
    e_off = [0]
    for e in Es[:-1]:
        e_off.append(e_off[-1] + e)
    host = torch.tensor([Es, e_off], dtype=torch.int64)
    if dev.type == "cuda":
        host = host.pin_memory()
    meta = host.to(dev, non_blocking=True)
    if E > 0:
        src_off = torch.repeat_interleave(torch.arange(B, device=dev, dtype=torch.int64) * n_src_each, meta[0], output_size=E)
        idx = idx + src_off
    rs = torch.stack([g["neighbors_row_splits"].to(torch.int64) for g in graphs])         # [B, m + 1], equal m per sample
    body = (rs[:, 1:] + meta[1][:, None]).reshape(-1)
    splits = torch.cat([rs.new_zeros(1), body])
    return GraphPlan(idx, splits, n_src_each * B)


# ------------------------------------------------------------------------------------------------
# dense
# ------------------------------------------------------------------------------------------------
def gemm(A, B, M, N, K, lda, ldb, transA, transB, out=None, bias=None, add=None, add_rows=0, act=ACT_NONE, pre=None,
         accumulate=False):
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=A.device)
    call("gaot_gemm_f32", ptr(A), ptr(B), ptr(out), M, N, K, lda, ldb, N, int(transA), int(transB), ptr(bias), ptr(add),
         int(add_rows), N, int(act), ptr(pre), int(accumulate))
    return out


def col_sum(x2d: torch.Tensor, M: int, N: int) -> torch.Tensor:
    out = torch.empty(N, dtype=torch.float32, device=x2d.device)
    call("gaot_col_sum", ptr(x2d), M, N, ptr(out), 0)
    return out


class _Linear(torch.autograd.Function):
    """y = act(x W^T + bias + add[row % add_rows]); x [M,K], W [N,K] (nn.Linear layout)."""

    @staticmethod
    def forward(ctx, x, W, bias, add, act):
        _need_cuda(x)
        x, W = _c(x), _c(W)
        M, K = x.shape
        N = W.shape[0]
        bias_c = _c(bias) if bias is not None else None
        add_c = _c(add) if add is not None else None
        add_rows = add_c.shape[0] if add_c is not None else 0
        if add_c is not None and (add_c.shape[1] != N or M % add_rows != 0):
            raise ValueError(f"broadcast-add of shape {tuple(add_c.shape)} does not tile an [{M},{N}] output")
        pre = torch.empty((M, N), dtype=torch.float32, device=x.device) if act != ACT_NONE else None
        y = gemm(x, W, M, N, K, K, K, 0, 1, bias=bias_c, add=add_c, add_rows=add_rows, act=act, pre=pre)
        ctx.save_for_backward(x, W, pre)
        ctx.act, ctx.has_bias, ctx.add_rows = act, bias is not None, add_rows
        return y

    @staticmethod
    def backward(ctx, dy):
        x, W, pre = ctx.saved_tensors
        dy = _c(dy)
        M, K = x.shape
        N = W.shape[0]
        if ctx.act != ACT_NONE:
            dpre = torch.empty_like(dy)
            call("gaot_act_bwd", ptr(dy), ptr(pre), ptr(dpre), dy.numel(), ctx.act)
        else:
            dpre = dy
        dx = dW = db = dadd = None
        if ctx.needs_input_grad[0]:
            dx = gemm(dpre, W, M, K, N, N, K, 0, 0)                  # [M,N] x [N,K]
        if ctx.needs_input_grad[1]:
            dW = gemm(dpre, x, N, K, M, N, K, 1, 0)                  # dpre^T [N,M] x [M,K]
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = col_sum(dpre, M, N)
        if ctx.add_rows and ctx.needs_input_grad[3]:
            dadd = dpre if ctx.add_rows == M else col_sum(dpre, M // ctx.add_rows, ctx.add_rows * N).view(ctx.add_rows, N)
        return dx, dW, db, dadd, None


def cast_bf16(x2d, want=True, want_t=False):
    """bf16 copy and/or bf16 transposed copy of an fp32 [rows, cols] matrix."""
    rows, cols = x2d.shape
    out = torch.empty((rows, cols), dtype=torch.bfloat16, device=x2d.device) if want else None
    out_t = torch.empty((cols, rows), dtype=torch.bfloat16, device=x2d.device) if want_t else None
    call("gaot_cast_bf16", ptr(x2d), ptr(out), rows, cols, ptr(out_t))
    return out, out_t


def transpose_bf16(xb):
    rows, cols = xb.shape
    out = torch.empty((cols, rows), dtype=torch.bfloat16, device=xb.device)
    call("gaot_transpose_bf16", ptr(xb), ptr(out), rows, cols)
    return out


def gemm_bf16(A, B, M, N, K, out_bf16=False, bias=None, add=None, add_rows=0, out=None, accumulate=False, act=ACT_NONE,
              pre=None):
    """C[M,N] = act(A[M,K] B[N,K]^T + bias + add) on the tensor cores (A, B bf16 row-major)."""
    if out is None:
        out = torch.empty((M, N), dtype=torch.bfloat16 if out_bf16 else torch.float32, device=A.device)
    call("gaot_gemm_bf16", ptr(A), ptr(B), ptr(out), M, N, K, int(out_bf16), ptr(bias), ptr(add), int(add_rows), N, int(act),
         ptr(pre), int(accumulate))
    return out


def gemm_bf16_nn(A, B, M, N, K, out_bf16=False):
    """C[M,N] = A[M,K] B[K,N] on the tensor cores (A, B bf16 row-major; B consumed MN-major)."""
    out = torch.empty((M, N), dtype=torch.bfloat16 if out_bf16 else torch.float32, device=A.device)
    call("gaot_gemm_bf16_nn", ptr(A), ptr(B), ptr(out), M, N, K, int(out_bf16))
    return out


def gemm_bf16_tn(A, B, M, N, R, out=None, accumulate=False):
    """C[M,N] (fp32) = A[R,M]^T B[R,N] on the tensor cores (A, B bf16 row-major, contraction over rows)."""
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=A.device)
    call("gaot_gemm_bf16_tn", ptr(A), ptr(B), ptr(out), M, N, R, int(accumulate))
    return out


# bf16 copies of weights prepared ahead of their GEMM (GAOT.forward does this on a side stream at the start of the step so
# the ~30 small cast / concatenation kernels run under the encoder instead of between the processor's GEMMs).  Keyed by the
# identity of the fp32 weight tensor object that the layer will pass to `linear`; the entry keeps that object alive.
_PREPARED = {}


def prepare_weight(W):
    """Register the tensor-core operand of W (2-D, fp32) - its bf16 copy, or in 'fp32' mode its fp16 hi / lo planes and
    scale - for the next `linear(x, W, ...)` call with this very object.  Returns the tensors it created (a caller that
    prepares on a side stream must record_stream them on the consuming stream)."""
    Wc = _c(W)
    if _PRECISION == "fp32":
        if not _split_ok(Wc.shape[0], Wc.shape[1]):
            return []
        ent = split_f16(Wc)
        _PREPARED[id(W)] = (W, ent)
        return list(ent)
    wb = cast_bf16(Wc)[0]
    _PREPARED[id(W)] = (W, wb)
    return [wb]


def clear_prepared():
    _PREPARED.clear()


# ---- split-precision (parity) tensor-core path -------------------------------------------------------------------------
SPLIT_WS_FLOATS = 1024                        # == gaot_split_ws_floats() (checked by tests/test_abi.py)


def split_f16(x2d):
    """(planes fp16 [2, rows, cols], scale fp32 [2] = {s, 1/s}) with x * s = hi + lo; s is a power of two set from max|x| on
    the device (no host read)."""
    rows, cols = x2d.shape
    planes = torch.empty((2, rows, cols), dtype=torch.float16, device=x2d.device)
    scale = torch.empty(2, dtype=torch.float32, device=x2d.device)
    ws = torch.empty(SPLIT_WS_FLOATS, dtype=torch.float32, device=x2d.device)
    call("gaot_split_f16", ptr(x2d), ptr(planes), ptr(scale), ptr(ws), rows * cols)
    return planes, scale


def _split_ok(N, K):
    return N % 8 == 0 and K % 8 == 0


def gemm_split(A, sa, B, sb, M, N, K, a_mn=False, b_mn=False, bias=None, add=None, add_rows=0, act=ACT_NONE, pre=None,
               out=None, accumulate=False):
    """C[M,N] fp32 = act(opA opB + bias + add) on split operands (see include/gaot_b200.h: gaot_gemm_split)."""
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=A.device)
    call("gaot_gemm_split", ptr(A), ptr(B), ptr(out), M, N, K, int(a_mn), int(b_mn), ptr(sa), ptr(sb), ptr(bias), ptr(add),
         int(add_rows), N, int(act), ptr(pre), int(accumulate))
    return out


class _LinearSplit(torch.autograd.Function):
    """y = act(x W^T + bias + add[row % add_rows]) in fp32-equivalent precision on tcgen05 (three-term fp16 split)."""

    @staticmethod
    def forward(ctx, x, W, bias, add, act, prepared):
        _need_cuda(x)
        x, W = _c(x), _c(W)
        M, K = x.shape
        N = W.shape[0]
        bias_c = _c(bias) if bias is not None else None
        add_c = _c(add) if add is not None else None
        add_rows = add_c.shape[0] if add_c is not None else 0
        if add_c is not None and (add_c.shape[1] != N or M % add_rows != 0):
            raise ValueError(f"broadcast-add of shape {tuple(add_c.shape)} does not tile an [{M},{N}] output")
        xs, sx = split_f16(x)
        ws, sw = prepared if prepared is not None else split_f16(W)
        pre = torch.empty((M, N), dtype=torch.float32, device=x.device) if act != ACT_NONE else None
        y = gemm_split(xs, sx, ws, sw, M, N, K, bias=bias_c, add=add_c, add_rows=add_rows, act=act, pre=pre)
        ctx.save_for_backward(xs, sx, ws, sw, pre)
        ctx.has_bias, ctx.add_rows, ctx.act = bias is not None, add_rows, act
        return y

    @staticmethod
    def backward(ctx, dy):
        xs, sx, ws, sw, pre = ctx.saved_tensors
        dy = _c(dy)
        M, K = xs.shape[1], xs.shape[2]
        N = ws.shape[1]
        if ctx.act != ACT_NONE:
            dpre = torch.empty_like(dy)
            call("gaot_act_bwd", ptr(dy), ptr(pre), ptr(dpre), dy.numel(), ctx.act)
            dy = dpre
        dx = dW = db = dadd = None
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            dys, sdy = split_f16(dy)
            if ctx.needs_input_grad[0]:
                dx = gemm_split(dys, sdy, ws, sw, M, K, N, b_mn=True)                 # dY W, W planes read MN-major
            if ctx.needs_input_grad[1]:
                dW = gemm_split(dys, sdy, xs, sx, N, K, M, a_mn=True, b_mn=True)      # dY^T X, both read MN-major in place
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = col_sum(dy, M, N)
        if ctx.add_rows and ctx.needs_input_grad[3]:
            dadd = dy if ctx.add_rows == M else col_sum(dy, M // ctx.add_rows, ctx.add_rows * N).view(ctx.add_rows, N)
        return dx, dW, db, dadd, None, None


class _LinearTC(torch.autograd.Function):
    """y = x W^T + bias + add[row % add_rows] with bf16 operands on tcgen05.  x may be fp32 (cast here) or
    already bf16; y is fp32, or bf16 when out_bf16 (feeding another tensor-core kernel)."""

    @staticmethod
    def forward(ctx, x, W, bias, add, out_bf16, act=ACT_NONE, prepared=None):
        _need_cuda(x)
        x = x if x.is_contiguous() else x.contiguous()
        W = _c(W)
        M, K = x.shape
        N = W.shape[0]
        bias_c = _c(bias) if bias is not None else None
        add_c = _c(add) if add is not None else None
        add_rows = add_c.shape[0] if add_c is not None else 0
        if add_c is not None and (add_c.shape[1] != N or M % add_rows != 0):
            raise ValueError(f"broadcast-add of shape {tuple(add_c.shape)} does not tile an [{M},{N}] output")
        xb = x if x.dtype == torch.bfloat16 else cast_bf16(_c(x))[0]
        wb = prepared if prepared is not None else cast_bf16(W)[0]
        pre = torch.empty((M, N), dtype=torch.float32, device=x.device) if act != ACT_NONE else None
        y = gemm_bf16(xb, wb, M, N, K, out_bf16=out_bf16, bias=bias_c, add=add_c, add_rows=add_rows, act=act, pre=pre)
        ctx.save_for_backward(xb, wb, pre)
        ctx.has_bias, ctx.add_rows, ctx.x_bf16, ctx.act = bias is not None, add_rows, x.dtype == torch.bfloat16, act
        return y

    @staticmethod
    def backward(ctx, dy):
        xb, wb, pre = ctx.saved_tensors
        dy = dy if dy.is_contiguous() else dy.contiguous()
        M, K = xb.shape
        N = wb.shape[0]
        dx = dW = db = dadd = None
        if ctx.act != ACT_NONE:
            if (M * N) % 4 == 0:
                dyb = torch.empty((M, N), dtype=torch.bfloat16, device=dy.device)
                call("gaot_act_bwd_bf16", ptr(dy), int(dy.dtype == torch.bfloat16), ptr(pre), ptr(dyb), M * N, ctx.act)
                dy = dyb                                        # from here on "dy" is the gradient w.r.t. the pre-activation
            else:
                dpre = torch.empty((M, N), dtype=torch.float32, device=dy.device)
                call("gaot_act_bwd", ptr(_c(dy)), ptr(pre), ptr(dpre), M * N, ctx.act)
                dy = dpre
        dyb = dy if dy.dtype == torch.bfloat16 else cast_bf16(dy)[0]
        if ctx.needs_input_grad[0]:
            if N % 8 == 0:
                dx = gemm_bf16_nn(dyb, wb, M, K, N, out_bf16=ctx.x_bf16)   # dY W, W read MN-major in its nn.Linear layout
            else:
                dx = gemm(dy.float(), wb.float(), M, K, N, N, K, 0, 0)
                dx = dx.to(torch.bfloat16) if ctx.x_bf16 else dx
        if ctx.needs_input_grad[1]:
            if N % 8 == 0:
                dW = gemm_bf16_tn(dyb, xb, N, K, M)                  # dY^T X, both operands read MN-major in place
            else:
                dW = gemm(dy.float(), xb.float(), N, K, M, N, K, 1, 0)
        if (ctx.has_bias and ctx.needs_input_grad[2]) or (ctx.add_rows and ctx.needs_input_grad[3]):
            dyf = dy if dy.dtype == torch.float32 else dy.float()
            if ctx.has_bias and ctx.needs_input_grad[2]:
                db = col_sum(dyf, M, N)
            if ctx.add_rows and ctx.needs_input_grad[3]:
                dadd = dyf if ctx.add_rows == M else col_sum(dyf, M // ctx.add_rows, ctx.add_rows * N).view(ctx.add_rows, N)
        return dx, dW, db, dadd, None, None, None


def linear(x, W, bias=None, add=None, act=ACT_NONE, tc=False, out_bf16=False):
    """tc=True marks a processor layer that may run on the tensor cores when precision is 'bf16';
    out_bf16 (tensor-core path only) keeps the result in bf16 for a following tensor-core kernel."""
    if tc and _PRECISION == "bf16" and x.shape[-1] % 8 == 0 and (act == ACT_NONE or W.shape[0] % 4 == 0):
        ent = _PREPARED.get(id(W))
        return _LinearTC.apply(x, W, bias, add, out_bf16, act, ent[1] if ent is not None else None)
    if tc and _PRECISION == "fp32" and _split_ok(W.shape[0], x.shape[-1]):
        ent = _PREPARED.get(id(W))
        return _LinearSplit.apply(x.float() if x.dtype != torch.float32 else x, W, bias, add, act,
                                  ent[1] if ent is not None else None)
    if x.dtype != torch.float32:
        x = x.float()
    return _Linear.apply(x, W, bias, add, act)


class _CondAffine(torch.autograd.Function):
    """y[b,t,:] = x[b,t,:] * scale[b,:] + bias[b,:]  (ConditionedNorm.forward, mlp.py:121-124)."""

    @staticmethod
    def forward(ctx, x, scale, bias):
        _need_cuda(x)
        x, scale, bias = _c(x.float()), _c(scale), _c(bias)
        B, T, D = x.shape
        if scale.shape != (B, D) or bias.shape != (B, D):
            raise ValueError(f"scale/bias must be [{B},{D}], got {tuple(scale.shape)} / {tuple(bias.shape)}")
        y = torch.empty_like(x)
        call("gaot_cond_affine_fwd", ptr(x), ptr(scale), ptr(bias), ptr(y), B, T, D)
        ctx.save_for_backward(x, scale)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, scale = ctx.saved_tensors
        B, T, D = x.shape
        dy = _c(dy.float())
        dx, ds, db = torch.empty_like(x), torch.empty_like(scale), torch.empty_like(scale)
        call("gaot_cond_affine_bwd", ptr(dy), ptr(x), ptr(scale), ptr(dx), ptr(ds), ptr(db), B, T, D)
        return dx, ds, db


def cond_affine(x, scale, bias):
    return _CondAffine.apply(x, scale, bias)


class _RMSNorm(torch.autograd.Function):
    """mode 'fp32' -> y; 'bf16' -> y16 (GEMM input only); 'both' -> (y, y16) (residual + GEMM input)."""

    @staticmethod
    def forward(ctx, x, gamma, eps, mode):
        x, gamma = _c(x), _c(gamma)
        rows, D = x.shape
        y = torch.empty_like(x) if mode in ("fp32", "both") else None
        y16 = torch.empty((rows, D), dtype=torch.bfloat16, device=x.device) if mode in ("bf16", "both") else None
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
        call("gaot_rmsnorm_fwd", ptr(x), ptr(gamma), ptr(y), ptr(y16), ptr(rstd), rows, D, float(eps))
        ctx.save_for_backward(x, gamma, rstd)
        ctx.mode = mode
        return y if mode == "fp32" else (y16 if mode == "bf16" else (y, y16))

    @staticmethod
    def backward(ctx, *grads):
        x, gamma, rstd = ctx.saved_tensors
        if ctx.mode == "fp32":
            dy, dy16 = grads[0], None
        elif ctx.mode == "bf16":
            dy, dy16 = None, grads[0]
        else:
            dy, dy16 = grads
        dy = _c(dy) if dy is not None else None
        if dy16 is not None:
            dy16 = dy16 if dy16.dtype == torch.bfloat16 else dy16.to(torch.bfloat16)
            dy16 = dy16 if dy16.is_contiguous() else dy16.contiguous()
        rows, D = x.shape
        dx = torch.empty_like(x)
        dg = torch.zeros(D, dtype=torch.float32, device=x.device)
        call("gaot_rmsnorm_bwd", ptr(dy), ptr(dy16), ptr(x), ptr(gamma), ptr(rstd), ptr(dx), ptr(dg), rows, D)
        return dx, dg, None, None


def rmsnorm(x, gamma, eps, mode="fp32"):
    return _RMSNorm.apply(x, gamma, eps, mode)


class _SwiGLU(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ab):
        ab = _c(ab)
        rows, F2 = ab.shape
        u = torch.empty((rows, F2 // 2), dtype=torch.float32, device=ab.device)
        call("gaot_swiglu_fwd", ptr(ab), ptr(u), rows, F2 // 2)
        ctx.save_for_backward(ab)
        return u

    @staticmethod
    def backward(ctx, du):
        (ab,) = ctx.saved_tensors
        du = _c(du)
        dab = torch.empty_like(ab)
        call("gaot_swiglu_bwd", ptr(du), ptr(ab), ptr(dab), ab.shape[0], ab.shape[1] // 2)
        return dab


class _SwiGLU16(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ab):
        ab = ab if ab.is_contiguous() else ab.contiguous()
        rows, F2 = ab.shape
        u = torch.empty((rows, F2 // 2), dtype=torch.bfloat16, device=ab.device)
        call("gaot_swiglu_fwd_bf16", ptr(ab), ptr(u), rows, F2 // 2)
        ctx.save_for_backward(ab)
        return u

    @staticmethod
    def backward(ctx, du):
        (ab,) = ctx.saved_tensors
        du = du if du.dtype == torch.bfloat16 else du.to(torch.bfloat16)
        du = du if du.is_contiguous() else du.contiguous()
        dab = torch.empty_like(ab)
        call("gaot_swiglu_bwd_bf16", ptr(du), ptr(ab), ptr(dab), ab.shape[0], ab.shape[1] // 2)
        return dab


# ---- gated FFN with the SwiGLU fused into the GEMM epilogues (bf16 mode) -------------------------------------------------
def ffn_fusable(D: int, F: int) -> bool:
    """Shapes the fused epilogues cover (gaot_gemm_bf16_swiglu / gaot_gemm_bf16_nn_swiglu_bwd): hidden width F in blocks
    of 32 and at least 128 (the two-chunk epilogue pass of the 128- and 256-column GEMM tiles)."""
    return _PRECISION == "bf16" and D % 8 == 0 and F % 32 == 0 and F >= 128 and not os.environ.get("GAOT_B200_NO_FFN_FUSE")


def interleave_w13(w1, w3):
    """[2F, D] weight of the fused w1 | w3 product with rows interleaved in blocks of 32 (w1[32j..], w3[32j..], ...), so an
    output tile holds a and b of the same features; differentiable (the gradient un-interleaves through autograd)."""
    F, D = w1.shape
    return torch.stack([w1.reshape(F // 32, 32, D), w3.reshape(F // 32, 32, D)], dim=1).reshape(2 * F, D)


class _SwiGLUFFN(torch.autograd.Function):
    """y = (silu(x w1^T) * (x w3^T)) w2^T + add  (reference FFN.forward, attn.py:137-146) as three tensor-core GEMMs:
    w13 with the SwiGLU in its epilogue (writes the pre-activations once, for the backward, and u), w2 with the residual
    in its epilogue; backward: dU = dY w2 is consumed inside its own epilogue (-> dA | dB), then dX and the two weight
    gradients.  No stand-alone SwiGLU pass, dU never reaches memory."""

    @staticmethod
    def forward(ctx, x, w13i, w2, add, prep13, prep2):
        _need_cuda(x)
        x = x if x.is_contiguous() else x.contiguous()
        M, D = x.shape
        F2, Dout = w13i.shape[0], w2.shape[0]
        xb = x if x.dtype == torch.bfloat16 else cast_bf16(_c(x))[0]
        w13b = prep13 if prep13 is not None else cast_bf16(_c(w13i))[0]
        w2b = prep2 if prep2 is not None else cast_bf16(_c(w2))[0]
        add_c = _c(add) if add is not None else None
        add_rows = add_c.shape[0] if add_c is not None else 0
        if add_c is not None and (add_c.shape[1] != Dout or M % add_rows != 0):
            raise ValueError(f"broadcast-add of shape {tuple(add_c.shape)} does not tile an [{M},{Dout}] output")
        ab = torch.empty((M, F2), dtype=torch.bfloat16, device=x.device)
        u = torch.empty((M, F2 // 2), dtype=torch.bfloat16, device=x.device)
        call("gaot_gemm_bf16_swiglu", ptr(xb), ptr(w13b), ptr(ab), ptr(u), M, F2, D)
        y = gemm_bf16(u, w2b, M, Dout, F2 // 2, bias=None, add=add_c, add_rows=add_rows)
        ctx.save_for_backward(xb, w13b, w2b, ab, u)
        ctx.add_rows, ctx.x_bf16 = add_rows, x.dtype == torch.bfloat16
        return y

    @staticmethod
    def backward(ctx, dy):
        xb, w13b, w2b, ab, u = ctx.saved_tensors
        M, D = xb.shape
        F2, Dout = w13b.shape[0], w2b.shape[0]
        dy = dy if dy.is_contiguous() else dy.contiguous()
        dyb = dy if dy.dtype == torch.bfloat16 else cast_bf16(dy)[0]
        dab = torch.empty_like(ab)
        call("gaot_gemm_bf16_nn_swiglu_bwd", ptr(dyb), ptr(w2b), ptr(ab), ptr(dab), M, F2 // 2, Dout)
        dW2 = gemm_bf16_tn(dyb, u, Dout, F2 // 2, M) if ctx.needs_input_grad[2] else None
        dx = gemm_bf16_nn(dab, w13b, M, D, F2, out_bf16=ctx.x_bf16) if ctx.needs_input_grad[0] else None
        dW13 = gemm_bf16_tn(dab, xb, F2, D, M) if ctx.needs_input_grad[1] else None
        dadd = None
        if ctx.add_rows and ctx.needs_input_grad[3]:
            dyf = dy if dy.dtype == torch.float32 else dy.float()
            dadd = dyf if ctx.add_rows == M else col_sum(dyf, M // ctx.add_rows, ctx.add_rows * Dout).view(ctx.add_rows, Dout)
        return dx, dW13, dW2, dadd, None, None


def swiglu_ffn(x, w13i, w2, add=None):
    """x [M, D] (bf16 or fp32), w13i = interleave_w13(w1, w3) [2F, D], w2 [Dout, F]: see _SwiGLUFFN."""
    e13, e2 = _PREPARED.get(id(w13i)), _PREPARED.get(id(w2))
    return _SwiGLUFFN.apply(x, w13i, w2, add, e13[1] if e13 is not None else None, e2[1] if e2 is not None else None)


def swiglu(ab):
    if ab.dtype == torch.bfloat16 and (ab.shape[1] // 2) % 8 == 0:
        return _SwiGLU16.apply(ab)
    return _SwiGLU.apply(ab.float() if ab.dtype != torch.float32 else ab)


class _Attention(torch.autograd.Function):
    """qkv [B*T, (H + 2*Hkv)*hd] = [q | k | v] per token -> o [B*T, H*hd]."""

    @staticmethod
    def forward(ctx, qkv, B, T, H, Hkv, hd):
        qkv = _c(qkv)
        ld = qkv.shape[1]
        o = torch.empty((B * T, H * hd), dtype=torch.float32, device=qkv.device)
        lse = torch.empty((B, H, T), dtype=torch.float32, device=qkv.device)
        scale = 1.0 / (hd ** 0.5)
        base = qkv.data_ptr()
        call("gaot_attention_fwd", base, base + 4 * H * hd, base + 4 * (H + Hkv) * hd, ptr(o), ptr(lse), B, T, H, Hkv, hd,
             ld, ld, scale)
        ctx.save_for_backward(qkv, o, lse)
        ctx.dims = (B, T, H, Hkv, hd, scale)
        return o

    @staticmethod
    def backward(ctx, do):
        qkv, o, lse = ctx.saved_tensors
        B, T, H, Hkv, hd, scale = ctx.dims
        do = _c(do)
        ld = qkv.shape[1]
        dqkv = torch.empty_like(qkv)
        delta = torch.empty_like(lse)
        b, d = qkv.data_ptr(), dqkv.data_ptr()
        ok, ov = 4 * H * hd, 4 * (H + Hkv) * hd
        call("gaot_attention_bwd", b, b + ok, b + ov, ptr(o), ptr(lse), ptr(do), ptr(delta), d, d + ok, d + ov, B, T, H, Hkv,
             hd, ld, ld, scale)
        return dqkv, None, None, None, None, None


class _AttentionTC(torch.autograd.Function):
    """Tensor-core attention: qkv bf16 [B*T, (H+2Hkv)*hd] -> o bf16 [B*T, H*hd]."""

    @staticmethod
    def forward(ctx, qkv, B, T, H, Hkv, hd):
        qkv = qkv if qkv.is_contiguous() else qkv.contiguous()
        ld = qkv.shape[1]
        o = torch.empty((B * T, H * hd), dtype=torch.bfloat16, device=qkv.device)
        lse = torch.empty((B, H, T), dtype=torch.float32, device=qkv.device)
        scale = 1.0 / (hd ** 0.5)
        call("gaot_attention_fwd_bf16", ptr(qkv), ptr(o), ptr(lse), B, T, H, Hkv, hd, ld, scale)
        ctx.save_for_backward(qkv, o, lse)
        ctx.dims = (B, T, H, Hkv, hd, scale)
        return o

    @staticmethod
    def backward(ctx, do):
        qkv, o, lse = ctx.saved_tensors
        B, T, H, Hkv, hd, scale = ctx.dims
        do = do if do.dtype == torch.bfloat16 else do.to(torch.bfloat16)
        do = do if do.is_contiguous() else do.contiguous()
        dqkv = torch.empty_like(qkv)
        delta = torch.empty_like(lse)
        dq_acc = torch.empty((B * T, H * hd), dtype=torch.float32, device=qkv.device)
        call("gaot_attention_bwd_bf16", ptr(qkv), ptr(o), ptr(lse), ptr(do), ptr(delta), ptr(dq_acc), ptr(dqkv), B, T, H, Hkv,
             hd, qkv.shape[1], scale)
        return dqkv, None, None, None, None, None


class _AttentionSplit(torch.autograd.Function):
    """Split-precision tensor-core attention: qkv fp32 [B*T, (H+2Hkv)*hd] -> o fp32 [B*T, H*hd] (hd 16 / 32)."""

    @staticmethod
    def forward(ctx, qkv, B, T, H, Hkv, hd):
        qkv = _c(qkv)
        ld = qkv.shape[1]
        qs, sq = split_f16(qkv)
        o = torch.empty((B * T, H * hd), dtype=torch.float32, device=qkv.device)
        lse = torch.empty((B, H, T), dtype=torch.float32, device=qkv.device)
        scale = 1.0 / (hd ** 0.5)
        call("gaot_attention_fwd_split", ptr(qs), ptr(sq), ptr(o), ptr(lse), B, T, H, Hkv, hd, ld, scale)
        ctx.save_for_backward(qs, sq, o, lse)
        ctx.dims = (B, T, H, Hkv, hd, scale, ld)
        return o

    @staticmethod
    def backward(ctx, do):
        qs, sq, o, lse = ctx.saved_tensors
        B, T, H, Hkv, hd, scale, ld = ctx.dims
        do = _c(do)
        dos, sg = split_f16(do)
        dqkv = torch.empty((B * T, ld), dtype=torch.float32, device=do.device)
        delta = torch.empty_like(lse)
        call("gaot_attention_bwd_split", ptr(qs), ptr(sq), ptr(o), ptr(lse), ptr(do), ptr(dos), ptr(sg), ptr(delta), ptr(dqkv),
             B, T, H, Hkv, hd, ld, scale)
        return dqkv, None, None, None, None, None


_ROPE_TABLES = {}


def rope_tables(T: int, hd: int, device):
    """cos / sin [T, hd] (fp32) of rotary-embedding-torch's default `RotaryEmbedding(dim=hd)` (reference attn.py:76, 106-108):
    angle(t, i) = t / 10000^(2i / hd), shared by the interleaved pair (2i, 2i + 1); cached per (T, hd, device)."""
    key = (T, hd, str(device))
    if key not in _ROPE_TABLES:
        inv = 1.0 / (10000 ** (torch.arange(0, hd, 2)[: hd // 2].float() / hd))
        ang = (torch.arange(T, dtype=torch.float32)[:, None] * inv[None, :]).repeat_interleave(2, dim=-1)
        _ROPE_TABLES[key] = (ang.cos().to(device), ang.sin().to(device))
    return _ROPE_TABLES[key]


def rope_qk(qkv, B, T, H, Hkv, hd):
    """Rotates the query and key heads of the fused qkv matrix [B*T, (H + 2 Hkv) * hd] by the token position (RoPE,
    positional_embedding='rope'); values pass through.  Plain differentiable torch ops in fp32 (an option outside the
    benchmarked configurations); the result keeps qkv's dtype and layout for the attention kernels."""
    cos, sin = rope_tables(T, hd, qkv.device)
    W = (H + Hkv) * hd
    x = qkv[:, :W].float().reshape(B, T, H + Hkv, hd)
    rot = torch.stack((-x[..., 1::2], x[..., 0::2]), dim=-1).reshape(x.shape)       # rotate_half on interleaved pairs
    y = x * cos[None, :, None, :] + rot * sin[None, :, None, :]
    return torch.cat([y.reshape(B * T, W).to(qkv.dtype), qkv[:, W:]], dim=1).contiguous()


def attention(qkv, B, T, H, Hkv, hd):
    if qkv.dtype == torch.bfloat16:
        return _AttentionTC.apply(qkv, B, T, H, Hkv, hd)
    if _PRECISION == "fp32" and hd in (16, 32) and qkv.shape[1] % 8 == 0 and qkv.shape[1] == (H + 2 * Hkv) * hd:
        return _AttentionSplit.apply(qkv, B, T, H, Hkv, hd)
    return _Attention.apply(qkv, B, T, H, Hkv, hd)


class _PatchPermute(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, grid, P, C, to_tokens):
        x = _c(x)
        B = x.shape[0]
        dims = len(grid)
        H, W = grid[0], grid[1]
        D = grid[2] if dims == 3 else 1
        nodes, pv = H * W * D, P ** dims
        out = torch.empty((B, nodes // pv, pv * C) if to_tokens else (B, nodes, C), dtype=torch.float32, device=x.device)
        call("gaot_patch_permute", ptr(x), ptr(out), B, dims, H, W, D, P, C, int(to_tokens))
        ctx.args = (grid, P, C, to_tokens)
        return out

    @staticmethod
    def backward(ctx, g):
        grid, P, C, to_tokens = ctx.args
        return _PatchPermute.apply(g, grid, P, C, not to_tokens), None, None, None, None


def patchify(x, grid, P):
    return _PatchPermute.apply(x, tuple(grid), P, x.shape[-1], True)


def unpatchify(x, grid, P, C):
    return _PatchPermute.apply(x, tuple(grid), P, C, False)


# ------------------------------------------------------------------------------------------------
# AGNO
# ------------------------------------------------------------------------------------------------
def edge_inputs(y, x, plan: GraphPlan):
    y, x = _c(y), _c(x)
    z = torch.empty((plan.E, y.shape[1] + x.shape[1]), dtype=torch.float32, device=y.device)
    call("gaot_edge_inputs", ptr(y), y.shape[1], ptr(x), x.shape[1], ptr(plan.index), ptr(plan.edge_row), plan.E, ptr(z))
    return z


def edge_diff(y, x, d, plan: GraphPlan):
    """z[e,:] = y[index[e], :d] - x[row(e), :d] (gemb.py:208-213); coordinates only, no gradient."""
    y, x = _c(y), _c(x)
    z = torch.empty((plan.E, d), dtype=torch.float32, device=y.device)
    call("gaot_edge_diff", ptr(y), y.shape[1], ptr(x), x.shape[1], int(d), ptr(plan.index), ptr(plan.edge_row), plan.E, ptr(z))
    return z


class _SegmentPool(torch.autograd.Function):
    """out[r,:] = reduce over the edges of row r of src[e,:] ('sum' | 'mean' | 'max', empty rows 0) with its backward;
    'max' routes the gradient to the first maximiser of each row and column (scatter_max, gemb.py:217)."""

    @staticmethod
    def forward(ctx, src, plan, reduce):
        code = {"sum": 0, "mean": 1, "max": 2}[reduce]
        src = _c(src)
        E, C = src.shape
        out = torch.empty((plan.m, C), dtype=torch.float32, device=src.device)
        call("gaot_segment_csr", ptr(src), ptr(plan.splits), ptr(out), 1, E, plan.m, C, code)
        ctx.save_for_backward(src, out)
        ctx.plan, ctx.code = plan, code
        return out

    @staticmethod
    def backward(ctx, g):
        src, out = ctx.saved_tensors
        plan = ctx.plan
        dsrc = torch.empty_like(src)
        call("gaot_segment_csr_bwd", ptr(_c(g)), ptr(src), ptr(out), ptr(plan.splits), ptr(plan.edge_row), ptr(dsrc), src.shape[0],
             plan.m, src.shape[1], ctx.code)
        return dsrc, None, None


def segment_pool(src, plan: GraphPlan, reduce: str):
    return _SegmentPool.apply(src, plan, reduce)


def edge_cosine_softmax(y, x, cd, plan: GraphPlan):
    y, x = _c(y), _c(x)
    alpha = torch.empty(max(plan.E, 1), dtype=torch.float32, device=y.device)[:plan.E]
    call("gaot_edge_cosine_softmax", ptr(y), y.shape[1], ptr(x), x.shape[1], int(cd), ptr(plan.index), ptr(plan.splits),
         plan.m, ptr(alpha))
    return alpha


class _EdgeDotSoftmax(torch.autograd.Function):
    """alpha[e] = softmax over the row of scale * <q[row(e)], k[index[e]]> (dot-product attention, agno.py:214-217);
    q [m, A], k [n, A] are the per-node projections, both differentiable."""

    @staticmethod
    def forward(ctx, q, k, plan, scale):
        q, k = _c(q), _c(k)
        alpha = torch.empty(max(plan.E, 1), dtype=torch.float32, device=q.device)[:plan.E]
        if plan.E > 0:                                     # a graph without any edge has no weights to compute
            call("gaot_edge_dot_softmax_fwd", ptr(q), ptr(k), q.shape[1], ptr(plan.index), ptr(plan.splits), plan.m, float(scale),
                 ptr(alpha))
        ctx.save_for_backward(q, k, alpha)
        ctx.plan, ctx.scale = plan, float(scale)
        return alpha

    @staticmethod
    def backward(ctx, dalpha):
        q, k, alpha = ctx.saved_tensors
        plan = ctx.plan
        if plan.E == 0:
            return torch.zeros_like(q), torch.zeros_like(k), None, None
        dq, dk = torch.empty_like(q), torch.empty_like(k)
        call("gaot_edge_dot_softmax_bwd", ptr(_c(dalpha)), ptr(alpha), ptr(q), ptr(k), q.shape[1], ptr(plan.index), ptr(plan.splits),
             plan.m, k.shape[0], ctx.scale, ptr(dq), ptr(dk))
        return dq, dk, None, None


def edge_dot_softmax(q, k, plan: GraphPlan, scale: float):
    return _EdgeDotSoftmax.apply(q, k, plan, scale)


def edge_mlp_supported(kin: int, hidden: int, n_hidden: int, C: int) -> bool:
    """Shapes the fused edge-weight kernels are instantiated for (== gaot_edge_mlp_supported)."""
    return kin in (4, 6, 32, 48) and hidden == 64 and n_hidden in (1, 2) and 16 <= C <= 64 and C % 16 == 0


class _EdgeMlp(torch.autograd.Function):
    """w[e,:] = alpha[e] * MLP(cat[y[index[e]], x[row(e)]]) in one kernel, forward and backward (gaot_edge_mlp_{fwd,bwd});
    params = (W1, b1, Wh0, bh0[, Wh1, bh1], Wo, bo).  alpha and the coordinates carry no gradient."""

    @staticmethod
    def forward(ctx, y, x, plan, alpha, *params):
        _need_cuda(y)
        y, x = _c(y), _c(x)
        params = [_c(p) for p in params]
        n_hidden = len(params) // 2 - 2
        W1, b1, Wo, bo = params[0], params[1], params[-2], params[-1]
        hid = params[2:-2]
        kc = W1.shape[1] // 2
        C = Wo.shape[0]
        if not edge_mlp_supported(W1.shape[1], W1.shape[0], n_hidden, C) or y.shape[1] < kc or x.shape[1] < kc:
            raise ValueError("edge_mlp: shape not instantiated in the fused kernels")
        alpha_c = _c(alpha) if alpha is not None else None
        w = torch.empty((plan.E, C), dtype=torch.float32, device=y.device)
        ph = [ptr(t) for t in hid] + [0, 0] * (2 - n_hidden)
        if plan.E > 0:
            call("gaot_edge_mlp_fwd", ptr(y), y.shape[1], ptr(x), x.shape[1], kc, ptr(plan.index), ptr(plan.edge_row), plan.E,
                 ptr(alpha_c), ptr(W1), ptr(b1), ph[0], ph[1], ph[2], ph[3], n_hidden, ptr(Wo), ptr(bo), C, ptr(w))
        ctx.save_for_backward(y, x, alpha_c if alpha_c is not None else y.new_zeros(0), *params)
        ctx.plan, ctx.has_alpha, ctx.n_hidden, ctx.kc, ctx.C = plan, alpha_c is not None, n_hidden, kc, C
        return w

    @staticmethod
    def backward(ctx, dw):
        y, x, alpha, *params = ctx.saved_tensors
        plan, n_hidden, kc, C = ctx.plan, ctx.n_hidden, ctx.kc, ctx.C
        dw = _c(dw)
        W1, b1, Wo, bo = params[0], params[1], params[-2], params[-1]
        hid = params[2:-2]
        grads = [torch.empty_like(p) for p in params]
        ph = [ptr(t) for t in hid] + [0, 0] * (2 - n_hidden)
        gh = [ptr(t) for t in grads[2:-2]] + [0, 0] * (2 - n_hidden)
        call("gaot_edge_mlp_bwd", ptr(y), y.shape[1], ptr(x), x.shape[1], kc, ptr(plan.index), ptr(plan.edge_row), plan.E,
             ptr(alpha) if ctx.has_alpha else 0, ptr(W1), ptr(b1), ph[0], ph[1], ph[2], ph[3], n_hidden, ptr(Wo), ptr(bo), C,
             ptr(dw), ptr(grads[0]), ptr(grads[1]), gh[0], gh[1], gh[2], gh[3], ptr(grads[-2]), ptr(grads[-1]))
        return (None, None, None, None, *grads)


def edge_mlp(y, x, plan: GraphPlan, alpha, params):
    """Fused edge weights (see _EdgeMlp); `params` = [W1, b1, Wh0, bh0, (Wh1, bh1,) Wo, bo]."""
    return _EdgeMlp.apply(y, x, plan, alpha, *params)


# ---- edge-sharded evaluation of the (batch-independent) edge weights under data parallelism --------------------------------
# Data parallelism replicates everything that does not depend on the sample - above all the edge-weight MLP over all E edges,
# forward and backward.  With a small per-GPU batch and a large mesh (BASELINE configs[4]: 262 k points, global batch 8) that
# replicated work IS the step.  Sharding it by DESTINATION ROW: rank r owns 1 / world of the rows of a MAGNO graph (latent
# tokens in the encoder, query points in the decoder) for ALL samples of the global batch - it evaluates the edge-weight MLP,
# the gather / reduce and their backward for the edges of its rows only.  Samples are exchanged twice per operator, on
# activations that are small next to the [E, C] edge tensors: an all-gather of the operator's input over the batch and an
# all-to-all that turns (all samples, my rows) into (my samples, all rows); the backward runs the inverse exchanges
# (reduce-scatter / all-to-all).  Parameter gradients produced this way are partial per rank and are completed by the gradient
# all-reduce of the step, which sums over ranks anyway.
SHARD_EDGES_MIN = 1 << 18          # edges from which sharding pays
_shard_group = None                # None: sharding off; set by enable_edge_sharding(process_group)


def enable_edge_sharding(group=True):
    """Switch the edge-sharded evaluation on (group=True: the default process group) or off (None / False)."""
    global _shard_group
    _shard_group = group if group else None


def edge_shard_world():
    import torch.distributed as dist
    if _shard_group is None or not dist.is_available() or not dist.is_initialized():
        return 1, 0, None
    grp = None if _shard_group is True else _shard_group
    return dist.get_world_size(grp), dist.get_rank(grp), grp


def _all_gather_rows(out, x, group):
    import torch.distributed as dist
    if dist.get_backend(group) == "nccl":
        dist.all_gather_into_tensor(out, x, group=group)
    else:                                                  # gloo (CPU tests of the host logic)
        dist.all_gather(list(out.chunk(dist.get_world_size(group), 0)), x, group=group)


def _reduce_scatter_rows(out, g, group):
    import torch.distributed as dist
    if dist.get_backend(group) == "nccl":
        dist.reduce_scatter_tensor(out, g, op=dist.ReduceOp.SUM, group=group)
    else:
        tmp = g.clone()
        dist.all_reduce(tmp, op=dist.ReduceOp.SUM, group=group)
        out.copy_(tmp.chunk(dist.get_world_size(group), 0)[dist.get_rank(group)])


class _AllGatherBatch(torch.autograd.Function):
    """x [Bl, ...] of every rank -> [world * Bl, ...] in rank order; backward: reduce-scatter (sum over ranks) of the gradient."""

    @staticmethod
    def forward(ctx, x, group):
        import torch.distributed as dist
        world = dist.get_world_size(group)
        out = torch.empty((world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        _all_gather_rows(out, x.contiguous(), group)
        ctx.group = group
        return out

    @staticmethod
    def backward(ctx, g):
        import torch.distributed as dist
        world = dist.get_world_size(ctx.group)
        out = torch.empty((g.shape[0] // world,) + tuple(g.shape[1:]), dtype=g.dtype, device=g.device)
        _reduce_scatter_rows(out, g.contiguous(), ctx.group)
        return out, None


def _exchange(x, group):
    import torch.distributed as dist
    out = torch.empty_like(x)
    dist.all_to_all_single(out, x, group=group)
    return out


class _RowsToSamples(torch.autograd.Function):
    """[world * Bl, rows_s, C] (ALL samples, MY destination rows) -> [Bl, world * rows_s, C] (MY samples, ALL rows): one
    all-to-all over NVLink; the backward is the inverse exchange."""

    @staticmethod
    def forward(ctx, x, group):
        import torch.distributed as dist
        world = dist.get_world_size(group)
        Bg, rows_s, C = x.shape
        Bl = Bg // world
        got = _exchange(x.contiguous(), group).view(world, Bl, rows_s, C)          # [source rank = row block, my samples, rows, C]
        ctx.group, ctx.dims = group, (world, Bl, rows_s, C)
        return got.permute(1, 0, 2, 3).reshape(Bl, world * rows_s, C)

    @staticmethod
    def backward(ctx, g):
        world, Bl, rows_s, C = ctx.dims
        send = g.view(Bl, world, rows_s, C).permute(1, 0, 2, 3).contiguous()         # [row block, my samples, rows, C]
        return _exchange(send.view(world * Bl, rows_s, C), ctx.group), None


def all_gather_batch(x, group):
    return _AllGatherBatch.apply(x, group)


def rows_to_samples(x, group):
    return _RowsToSamples.apply(x, group)


class _ScaleRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, k, alpha):
        k = _c(k)
        out = torch.empty_like(k)
        call("gaot_scale_rows", ptr(k), ptr(alpha), k.shape[0], k.shape[1], ptr(out))
        ctx.save_for_backward(alpha, k if alpha.requires_grad else None)
        return out

    @staticmethod
    def backward(ctx, g):
        alpha, k = ctx.saved_tensors
        g = _c(g)
        out = torch.empty_like(g)
        call("gaot_scale_rows", ptr(g), ptr(alpha), g.shape[0], g.shape[1], ptr(out))
        dalpha = None
        if k is not None and ctx.needs_input_grad[1]:      # learned attention (dot-product): dalpha[e] = <g[e,:], kappa[e,:]>
            dalpha = torch.empty(max(g.shape[0], 1), dtype=torch.float32, device=g.device)[:g.shape[0]]
            call("gaot_row_dot", ptr(g), ptr(k), g.shape[0], g.shape[1], ptr(dalpha))
        return out, dalpha


def scale_rows(k, alpha):
    return _ScaleRows.apply(k, alpha)


class _Aggregate(torch.autograd.Function):
    """out[b,r,:] = sum_{e in row r} w[e,:] * f[b,index[e],:]  (agno.py:198,245-271)."""

    @staticmethod
    def forward(ctx, f, w, plan):
        f, w = _c(f), _c(w)
        B, Ns, C = f.shape
        if Ns != plan.n_src or w.shape[0] != plan.E or w.shape[1] != C:
            raise ValueError(f"aggregate: f {tuple(f.shape)} / w {tuple(w.shape)} do not match graph (n_src={plan.n_src}, E={plan.E})")
        out = torch.empty((B, plan.m, C), dtype=torch.float32, device=f.device)
        call("gaot_agno_aggregate", ptr(f), ptr(w), 0, ptr(plan.index), ptr(plan.splits), ptr(out), B, Ns, plan.m, C)
        ctx.save_for_backward(f, w)
        ctx.plan = plan
        return out

    @staticmethod
    def backward(ctx, g):
        f, w = ctx.saved_tensors
        plan = ctx.plan
        g = _c(g)
        B, Ns, C = f.shape
        df = dw = None
        if ctx.needs_input_grad[0]:
            ts, tr, tp = plan.transposed()
            df = torch.empty_like(f)
            call("gaot_agno_aggregate", ptr(g), ptr(w), ptr(tp), ptr(tr), ptr(ts), ptr(df), B, plan.m, Ns, C)
        if ctx.needs_input_grad[1]:
            dw = torch.empty_like(w)
            call("gaot_agno_weight_grad", ptr(g), ptr(f), ptr(plan.index), ptr(plan.splits), ptr(dw), B, Ns, plan.m, plan.E, C)
        return df, dw, None


def agno_aggregate(f, w, plan):
    return _Aggregate.apply(f, w, plan)


FUSE_MAX_CIN, FUSE_MAX_COUT, FUSE_MAX_C = 4, 4, 128


class _LiftAggregate(torch.autograd.Function):
    """out[b,r,:] = sum_{e in r} w[e,:] * (Wl pn[b,index[e],:] + bl): lifting fused into the gather."""

    @staticmethod
    def forward(ctx, pn, Wl, bl, w, plan):
        pn, Wl, bl, w = _c(pn), _c(Wl), _c(bl), _c(w)
        B, Ns, Cin = pn.shape
        C = Wl.shape[0]
        if Ns != plan.n_src or w.shape[0] != plan.E or w.shape[1] != C:
            raise ValueError(f"lift_aggregate: pn {tuple(pn.shape)} / w {tuple(w.shape)} do not match graph")
        out = torch.empty((B, plan.m, C), dtype=torch.float32, device=pn.device)
        # sample-minor copy [Ns, B, Cin] (1 MB at cfg-2): the kernels read the 8 samples of a work item with two 128-bit
        # loads per edge instead of 8 scalar gathers (used when B is a multiple of the 8-sample chunk)
        pn_t = pn.permute(1, 0, 2).contiguous() if (B % 8 == 0 and (8 * Cin) % 4 == 0) else None
        call("gaot_lift_aggregate_fwd", ptr(pn), ptr(Wl), ptr(bl), ptr(w), ptr(plan.index), ptr(plan.splits), ptr(out), B, Ns,
             plan.m, C, Cin, ptr(pn_t))
        ctx.save_for_backward(pn, Wl, bl, w, pn_t)
        ctx.plan = plan
        return out

    @staticmethod
    def backward(ctx, g):
        pn, Wl, bl, w, pn_t = ctx.saved_tensors
        plan = ctx.plan
        if ctx.needs_input_grad[0]:
            raise _lib.GaotLibraryError("lift_aggregate has no gradient for its raw inputs; use the unfused lifting + "
                                        "agno_aggregate path (MAGNOEncoder does when pndata.requires_grad)")
        g = _c(g)
        B, Ns, Cin = pn.shape
        C = Wl.shape[0]
        dw, dWl, dbl = torch.empty_like(w), torch.empty_like(Wl), torch.empty_like(bl)
        call("gaot_lift_aggregate_bwd", ptr(g), ptr(pn), ptr(Wl), ptr(bl), ptr(w), ptr(plan.index), ptr(plan.splits), ptr(dw),
             ptr(dWl), ptr(dbl), B, Ns, plan.m, plan.E, C, Cin, ptr(pn_t))
        return None, dWl, dbl, dw, None


def lift_aggregate(pn, Wl, bl, w, plan):
    return _LiftAggregate.apply(pn, Wl, bl, w, plan)


# Decoder aggregate: walk the graph by source node (default; reads every latent row once per sample) or by destination row
# (deterministic summation order; GAOT_B200_ROW_WALK=1 or ops.ROW_WALK = True).
import os as _os
ROW_WALK = bool(int(_os.environ.get("GAOT_B200_ROW_WALK", "0")))


def source_walk_pays(B, C):
    """The lane-per-sample source-walk kernels need a batch that fills most of a warp and a latent row that fits a lane's
    registers; small batches (vx mode's merged graphs, B = 1 per GPU) stay on the row-walk kernels."""
    return B >= 16 and C % 16 == 0 and C <= 64


class _AgnoDot(torch.autograd.Function):
    """out[b,r,o] = sum_{e in r} <u[o,e,:], f[b,index[e],:]>: aggregate fused with an affine read-out."""

    @staticmethod
    def forward(ctx, f, u, plan):
        f, u = _c(f), _c(u)
        B, Ns, C = f.shape
        Cout = u.shape[0]
        if Ns != plan.n_src or u.shape[1] != plan.E or u.shape[2] != C:
            raise ValueError(f"agno_dot: f {tuple(f.shape)} / u {tuple(u.shape)} do not match graph")
        out = torch.empty((B, plan.m, Cout), dtype=torch.float32, device=f.device)
        ctx.row_walk = ROW_WALK or not source_walk_pays(B, C)
        if ctx.row_walk:
            call("gaot_agno_dot_fwd", ptr(f), ptr(u), ptr(plan.index), ptr(plan.splits), ptr(out), B, Ns, plan.m, plan.E, C, Cout)
        else:
            ts, tr, tp = plan.transposed()
            call("gaot_agno_dot_fwd_t", ptr(f), ptr(u), ptr(ts), ptr(tr), ptr(tp), ptr(out), B, Ns, plan.m, plan.E, C, Cout)
        ctx.save_for_backward(f, u)
        ctx.plan = plan
        return out

    @staticmethod
    def backward(ctx, g):
        f, u = ctx.saved_tensors
        plan = ctx.plan
        g = _c(g)
        B, Ns, C = f.shape
        Cout = u.shape[0]
        df = torch.empty_like(f) if ctx.needs_input_grad[0] else None
        du = torch.empty_like(u) if ctx.needs_input_grad[1] else None
        if df is None and du is None:
            return None, None, None
        if ctx.row_walk:
            ts, tr, tp = plan.transposed() if df is not None else (None, None, None)
            call("gaot_agno_dot_bwd", ptr(g), ptr(f), ptr(u), ptr(plan.index), ptr(plan.splits), ptr(ts), ptr(tr), ptr(tp), ptr(df),
                 ptr(du), B, Ns, plan.m, plan.E, C, Cout)
        else:
            ts, tr, tp = plan.transposed()
            call("gaot_agno_dot_bwd_t", ptr(g), ptr(f), ptr(u), ptr(plan.index), ptr(plan.splits), ptr(ts), ptr(tr), ptr(tp), ptr(df),
                 ptr(du), B, Ns, plan.m, plan.E, C, Cout)
        return df, du, None


def agno_dot(f, u, plan):
    return _AgnoDot.apply(f, u, plan)


def segment_csr(src: torch.Tensor, indptr: torch.Tensor, reduce: str = "sum", use_scatter: bool = True) -> torch.Tensor:
    """Drop-in for reference segment_csr (src/model/layers/utils/segment_csr.py:14-19), forward only:
    src [E,C] / [B,E,C] / [E], indptr [m+1] (or [B,m+1] repeated rows, as agno.py:266-269 builds it)."""
    _need_cuda(src)
    code = {"sum": 0, "mean": 1, "max": 2}.get(reduce)
    if code is None:
        raise ValueError("reduce must be one of 'mean', 'sum', 'max'")
    if indptr.ndim == 2:
        indptr = indptr[0]
    s = _c(src)
    shape = s.shape
    if s.ndim == 1:
        s3 = s.view(1, -1, 1)
    elif s.ndim == 2:
        s3 = s.view(1, shape[0], shape[1])
    else:
        s3 = s
    B, E, C = s3.shape
    m = indptr.numel() - 1
    out = torch.empty((B, m, C), dtype=torch.float32, device=s.device)
    ip = indptr.to(torch.int64).contiguous()
    call("gaot_segment_csr", ptr(s3), ptr(ip), ptr(out), B, E, m, C, code)
    if src.ndim == 1:
        return out.view(m)
    if src.ndim == 2:
        return out.view(m, C)
    return out


def geo_stats(src, queries, plan: GraphPlan, groups: int = 1):
    """Standardised statistical features [m, 3+2d] (gemb.py:83-171); `groups` consecutive row blocks
    are standardised separately (vx mode merged graphs)."""
    src, queries = _c(src), _c(queries)
    d = queries.shape[1]
    feat = torch.empty((plan.m, 3 + 2 * d), dtype=torch.float32, device=src.device)
    call("gaot_geo_stats", ptr(src), ptr(queries), d, ptr(plan.index), ptr(plan.splits), plan.m, ptr(feat))
    call("gaot_col_standardize", ptr(feat), plan.m // groups, 3 + 2 * d, groups)
    return feat


# ------------------------------------------------------------------------------------------------
# step
# ------------------------------------------------------------------------------------------------
class _MSE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target):
        pred, target = _c(pred), _c(target)
        loss = torch.zeros(1, dtype=torch.float32, device=pred.device)
        dpred = torch.empty_like(pred)
        call("gaot_mse_fwd_bwd", ptr(pred), ptr(target), ptr(loss), ptr(dpred), pred.numel(), 1.0)
        ctx.save_for_backward(dpred)
        return loss.view(())

    @staticmethod
    def backward(ctx, g):
        (dpred,) = ctx.saved_tensors
        return dpred * g, None


def mse_loss(pred, target):
    return _MSE.apply(pred, target)


def adamw_step(param, grad, exp_avg, exp_avg_sq, step, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, grad_scale=1.0):
    call("gaot_adamw_step", ptr(param), ptr(grad), ptr(exp_avg), ptr(exp_avg_sq), param.numel(), float(lr), float(betas[0]),
         float(betas[1]), float(eps), float(weight_decay), int(step), float(grad_scale))


def rollout_update(pred, u_in, dt, mode, u_mean, u_std, s_mean, s_std):
    """(u_phys, u_next): the de-normalised prediction of the step (what the reference returns, gaot.py:425-430, 436-476)
    and the re-normalised state carried into the next step (gaot.py:433)."""
    pred = _c(pred)
    u_in = _c(u_in) if u_in is not None else None
    C = pred.shape[-1]
    phys, nxt = torch.empty_like(pred), torch.empty_like(pred)
    call("gaot_rollout_update", ptr(pred), ptr(u_in), ptr(phys), ptr(nxt), pred.numel() // C, C,
         float(dt), int(mode), ptr(u_mean), ptr(u_std), ptr(s_mean), ptr(s_std))
    return phys, nxt
