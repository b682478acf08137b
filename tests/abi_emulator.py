"""TEST INFRASTRUCTURE ONLY: a CPU stand-in for libgaot_b200.so's entry points, used by the
`-m "not gpu"` tests to exercise the *host logic* (argument order, shapes, autograd formulas, graph
merging, flat-buffer optimizer, gloo data parallelism) in a container without a GPU.

`install()` monkeypatches `gaot_b200._lib.call` for the duration of a test; the product never imports
this file and has no CPU path of its own (`gaot_b200._lib.call` on CPU tensors raises).  Each emulated
function receives exactly what the C function would: raw addresses and sizes, in the order
include/gaot_b200.h declares them; it views host memory through ctypes/numpy and computes with
numpy/torch in fp64 where that is cheap.
"""
from __future__ import annotations

import contextlib
import ctypes
import math

import numpy as np
import torch

from oracle import gaot_oracle as orc

_CT = {np.float32: ctypes.c_float, np.int64: ctypes.c_int64}


def _arr(ptr, shape, dtype=np.float32):
    n = int(np.prod(shape)) if len(shape) else 1
    if n == 0 or not ptr:
        return np.zeros(shape, dtype=dtype)
    return np.ctypeslib.as_array(ctypes.cast(int(ptr), ctypes.POINTER(_CT[dtype])), shape=(n,)).reshape(shape)


def _t(ptr, shape, dtype=np.float32):
    return torch.from_numpy(_arr(ptr, shape, dtype))


def _bf(ptr, shape):
    """bf16 host array viewed as a torch.bfloat16 tensor"""
    n = int(np.prod(shape))
    raw = np.ctypeslib.as_array(ctypes.cast(int(ptr), ctypes.POINTER(ctypes.c_uint16)), shape=(n,))
    return torch.from_numpy(raw.view(np.int16)).view(torch.bfloat16).view(shape)


def _csr(index_ptr, splits_ptr, m):
    rs = _t(splits_ptr, (m + 1,), np.int64)
    E = int(rs[-1])
    return {"neighbors_index": _t(index_ptr, (E,), np.int64), "neighbors_row_splits": rs}, E


def _act(x, act):
    return torch.nn.functional.gelu(x) if act == 1 else (torch.relu(x) if act == 2 else x)


# ---- graph -------------------------------------------------------------------------------------
def gaot_radius_count(data, n, queries, m, d, radius, inclusive, cap, batch, db, qb, row_splits):
    D = _t(data, (batch if db else 1, n, d))
    Q = _t(queries, (batch if qb else 1, m, d))
    rs = _arr(row_splits, (batch, m + 1), np.int64)
    for b in range(batch):
        g = orc.radius_csr(D[b if db else 0], Q[b if qb else 0], radius, bool(inclusive), cap if cap > 0 else None)
        rs[b] = g["neighbors_row_splits"].numpy()


def gaot_radius_fill(data, n, queries, m, d, radius, inclusive, cap, batch, db, qb, row_splits, offsets, index):
    D = _t(data, (batch if db else 1, n, d))
    Q = _t(queries, (batch if qb else 1, m, d))
    offs = _arr(offsets, (batch,), np.int64)
    for b in range(batch):
        g = orc.radius_csr(D[b if db else 0], Q[b if qb else 0], radius, bool(inclusive), cap if cap > 0 else None)
        e = g["neighbors_index"].numel()
        _arr(index + 8 * int(offs[b]), (e,), np.int64)[:] = g["neighbors_index"].numpy()


def gaot_grid_bin(data, n, d, batch, lox, loy, loz, h, gx, gy, gz, cell_start, srt, cell_ws, cursor_ws):
    pass        # the emulated search below does not need the cell structure


def gaot_radius_count_grid(data, n, queries, m, d, radius, inclusive, batch, db, qb, lox, loy, loz, h, gx, gy, gz, cell_start, srt,
                           row_splits):
    assert h >= radius * 1.0005
    gaot_radius_count(data, n, queries, m, d, radius, inclusive, 0, batch, db, qb, row_splits)


def gaot_radius_fill_grid(data, n, queries, m, d, radius, inclusive, batch, db, qb, lox, loy, loz, h, gx, gy, gz, cell_start, srt,
                          row_splits, offsets, index):
    gaot_radius_fill(data, n, queries, m, d, radius, inclusive, 0, batch, db, qb, row_splits, offsets, index)


def gaot_csr_edge_rows(row_splits, m, edge_row):
    rs = _arr(row_splits, (m + 1,), np.int64)
    _arr(edge_row, (int(rs[-1]),), np.int64)[:] = np.repeat(np.arange(m, dtype=np.int64), np.diff(rs))


def gaot_csr_transpose(index, row_splits, m, n_src, E, t_splits, t_rows, t_perm, ws):
    g, _ = _csr(index, row_splits, m)
    t = orc.csr_transpose(g, n_src)
    _arr(t_splits, (n_src + 1,), np.int64)[:] = t["neighbors_row_splits"].numpy()
    _arr(t_rows, (E,), np.int64)[:] = t["neighbors_index"].numpy()
    _arr(t_perm, (E,), np.int64)[:] = t["edge_perm"].numpy()


# ---- AGNO ---------------------------------------------------------------------------------------
def gaot_edge_inputs(y, dy, x, dx, index, edge_row, E, z):
    idx, row = _arr(index, (E,), np.int64), _arr(edge_row, (E,), np.int64)
    n_y, n_x = (int(idx.max()) + 1 if E else 0), (int(row.max()) + 1 if E else 0)
    _arr(z, (E, dy + dx))[:] = np.concatenate([_arr(y, (n_y, dy))[idx], _arr(x, (n_x, dx))[row]], axis=1)


def gaot_edge_diff(y, dy, x, dx, d, index, edge_row, E, z):
    idx, row = _arr(index, (E,), np.int64), _arr(edge_row, (E,), np.int64)
    n_y, n_x = (int(idx.max()) + 1 if E else 0), (int(row.max()) + 1 if E else 0)
    _arr(z, (E, d))[:] = _arr(y, (n_y, dy))[idx][:, :d] - _arr(x, (n_x, dx))[row][:, :d]


def gaot_segment_csr_bwd(dout, src, out, row_splits, edge_row, dsrc, E, m, C, reduce):
    rs, row = _arr(row_splits, (m + 1,), np.int64), _arr(edge_row, (E,), np.int64)
    g = _arr(dout, (m, C))[row].astype(np.float64)
    if reduce == 1:
        g = g / (rs[1:] - rs[:-1])[row][:, None]
    elif reduce == 2:
        S, O = _arr(src, (E, C)), _arr(out, (m, C))
        hit = S == O[row]
        first = np.zeros_like(hit)
        for r in range(m):                                   # first maximiser per row and column
            a, e = int(rs[r]), int(rs[r + 1])
            if e > a:
                h = hit[a:e]
                pos = h.argmax(axis=0)
                first[a + pos, np.arange(C)] = h[pos, np.arange(C)]
        g = g * first
    _arr(dsrc, (E, C))[:] = g.astype(np.float32)


def gaot_edge_cosine_softmax(y, dy, x, dx, cd, index, row_splits, m, alpha):
    g, E = _csr(index, row_splits, m)
    idx = g["neighbors_index"]
    seg, _ = orc._seg_ids(g["neighbors_row_splits"])
    n_y = int(idx.max()) + 1 if E else 0
    yy, xx = _t(y, (n_y, dy)).double()[:, :cd], _t(x, (m, dx)).double()[:, :cd]
    s = (torch.nn.functional.normalize(xx[seg], dim=-1) * torch.nn.functional.normalize(yy[idx], dim=-1)).sum(-1)
    ex = torch.exp(s - orc.segment_csr(s, g["neighbors_row_splits"], "max")[seg])
    _arr(alpha, (E,))[:] = (ex / orc.segment_csr(ex, g["neighbors_row_splits"], "sum")[seg]).float().numpy()


def gaot_edge_dot_softmax_fwd(q, k, A, index, row_splits, m, scale, alpha):
    g, E = _csr(index, row_splits, m)
    idx = g["neighbors_index"]
    seg, _ = orc._seg_ids(g["neighbors_row_splits"])
    n = int(idx.max()) + 1 if E else 0
    Q, K = _t(q, (m, A)).double(), _t(k, (n, A)).double()
    s = (Q[seg] * K[idx]).sum(-1) * scale
    ex = torch.exp(s - orc.segment_csr(s, g["neighbors_row_splits"], "max")[seg])
    _arr(alpha, (E,))[:] = (ex / orc.segment_csr(ex, g["neighbors_row_splits"], "sum")[seg]).float().numpy()


def gaot_edge_dot_softmax_bwd(dalpha, alpha, q, k, A, index, row_splits, m, n, scale, dq, dk):
    g, E = _csr(index, row_splits, m)
    idx = g["neighbors_index"]
    seg, _ = orc._seg_ids(g["neighbors_row_splits"])
    Q, K = _t(q, (m, A)).double(), _t(k, (n, A)).double()
    al, da = _t(alpha, (E,)).double(), _t(dalpha, (E,)).double()
    t = orc.segment_csr(al * da, g["neighbors_row_splits"], "sum")[seg]
    ds = al * (da - t) * scale
    DQ = torch.zeros(m, A, dtype=torch.float64).index_add_(0, seg, ds[:, None] * K[idx])
    DK = torch.zeros(n, A, dtype=torch.float64).index_add_(0, idx, ds[:, None] * Q[seg])
    _arr(dq, (m, A))[:] = DQ.float().numpy()
    _arr(dk, (n, A))[:] = DK.float().numpy()


def gaot_row_dot(a, b, rows, C, out):
    _arr(out, (rows,))[:] = (_arr(a, (rows, C)).astype(np.float64) * _arr(b, (rows, C))).sum(-1).astype(np.float32)


def gaot_scale_rows(inp, scale, rows, C, out):
    _arr(out, (rows, C))[:] = _arr(inp, (rows, C)) * _arr(scale, (rows,))[:, None]


def gaot_agno_aggregate(f, w, w_perm, index, row_splits, out, B, Ns, Nd, C):
    g, E = _csr(index, row_splits, Nd)
    W = _t(w, (E, C)).double()
    if w_perm:
        W = W[_t(w_perm, (E,), np.int64)]
    msg = W.unsqueeze(0) * _t(f, (B, Ns, C)).double()[:, g["neighbors_index"], :]
    _arr(out, (B, Nd, C))[:] = orc.segment_csr(msg, g["neighbors_row_splits"], "sum").float().numpy()


def gaot_agno_weight_grad(gr, f, index, row_splits, dw, B, Ns, Nd, E, C):
    g, E2 = _csr(index, row_splits, Nd)
    assert E2 == E
    seg, _ = orc._seg_ids(g["neighbors_row_splits"])
    G, F = _t(gr, (B, Nd, C)).double(), _t(f, (B, Ns, C)).double()
    _arr(dw, (E, C))[:] = (G[:, seg, :] * F[:, g["neighbors_index"], :]).sum(0).float().numpy()


def gaot_lift_aggregate_fwd(pn, Wl, bl, w, index, row_splits, out, B, Ns, Nd, C, Cin, pn_t):
    if pn_t:                                              # the sample-minor copy must hold the same values
        assert torch.equal(_t(pn_t, (Ns, B, Cin)).permute(1, 0, 2), _t(pn, (B, Ns, Cin)))
    g, E = _csr(index, row_splits, Nd)
    f = _t(pn, (B, Ns, Cin)).double() @ _t(Wl, (C, Cin)).double().T + _t(bl, (C,)).double()
    msg = _t(w, (E, C)).double().unsqueeze(0) * f[:, g["neighbors_index"], :]
    _arr(out, (B, Nd, C))[:] = orc.segment_csr(msg, g["neighbors_row_splits"], "sum").float().numpy()


def gaot_lift_aggregate_bwd(gr, pn, Wl, bl, w, index, row_splits, dw, dWl, dbl, B, Ns, Nd, E, C, Cin, pn_t):
    if pn_t:
        assert torch.equal(_t(pn_t, (Ns, B, Cin)).permute(1, 0, 2), _t(pn, (B, Ns, Cin)))
    g, _ = _csr(index, row_splits, Nd)
    with torch.enable_grad():
        wl = _t(Wl, (C, Cin)).double().requires_grad_(True)
        b = _t(bl, (C,)).double().requires_grad_(True)
        ww = _t(w, (E, C)).double().requires_grad_(True)
        f = _t(pn, (B, Ns, Cin)).double() @ wl.T + b
        out = orc.segment_csr(ww.unsqueeze(0) * f[:, g["neighbors_index"], :], g["neighbors_row_splits"], "sum")
        out.backward(_t(gr, (B, Nd, C)).double())
    _arr(dw, (E, C))[:] = ww.grad.float().numpy()
    _arr(dWl, (C, Cin))[:] = wl.grad.float().numpy()
    _arr(dbl, (C,))[:] = b.grad.float().numpy()


def _dot_ref(f, u, g):
    msg = torch.einsum("oec,bec->beo", u, f[:, g["neighbors_index"], :])
    return orc.segment_csr(msg, g["neighbors_row_splits"], "sum")


def gaot_agno_dot_fwd(f, u, index, row_splits, out, B, Ns, Nd, E, C, Cout):
    g, _ = _csr(index, row_splits, Nd)
    _arr(out, (B, Nd, Cout))[:] = _dot_ref(_t(f, (B, Ns, C)).double(), _t(u, (Cout, E, C)).double(), g).float().numpy()


def gaot_agno_dot_bwd(gr, f, u, index, row_splits, ts, tr, tp, df, du, B, Ns, Nd, E, C, Cout):
    g, _ = _csr(index, row_splits, Nd)
    with torch.enable_grad():
        F_ = _t(f, (B, Ns, C)).double().requires_grad_(True)
        U = _t(u, (Cout, E, C)).double().requires_grad_(True)
        _dot_ref(F_, U, g).backward(_t(gr, (B, Nd, Cout)).double())
    if df:
        t = orc.csr_transpose(g, Ns)                      # the kernel walks the transposed graph it is handed
        assert np.array_equal(_arr(ts, (Ns + 1,), np.int64), t["neighbors_row_splits"].numpy())
        assert np.array_equal(_arr(tp, (E,), np.int64), t["edge_perm"].numpy())
        _arr(df, (B, Ns, C))[:] = F_.grad.float().numpy()
    if du:
        _arr(du, (Cout, E, C))[:] = U.grad.float().numpy()


def gaot_agno_dot_fwd_t(f, u, ts, tr, tp, out, B, Ns, Nd, E, C, Cout):
    F, U = _t(f, (B, Ns, C)).double(), _t(u, (Cout, E, C)).double()
    tsp, rows, perm = _arr(ts, (Ns + 1,), np.int64), _arr(tr, (E,), np.int64), _arr(tp, (E,), np.int64)
    src = np.repeat(np.arange(Ns, dtype=np.int64), np.diff(tsp))          # source node of transposed slot k
    o = torch.zeros((B, Nd, Cout), dtype=torch.float64)
    contrib = torch.einsum("oec,bec->beo", U[:, perm, :], F[:, src, :])   # [B, E, Cout] in transposed slot order
    o.index_add_(1, torch.from_numpy(rows), contrib)
    _t(out, (B, Nd, Cout))[:] = o.float()


def gaot_agno_dot_bwd_t(gr, f, u, index, row_splits, ts, tr, tp, df, du, B, Ns, Nd, E, C, Cout):
    G, F, U = _t(gr, (B, Nd, Cout)).double(), _t(f, (B, Ns, C)).double(), _t(u, (Cout, E, C)).double()
    tsp, rows, perm = _arr(ts, (Ns + 1,), np.int64), _arr(tr, (E,), np.int64), _arr(tp, (E,), np.int64)
    src = torch.from_numpy(np.repeat(np.arange(Ns, dtype=np.int64), np.diff(tsp)))
    gk = G[:, torch.from_numpy(rows), :]                                   # [B, E, Cout] in transposed slot order
    if df:
        d = torch.zeros((B, Ns, C), dtype=torch.float64)
        d.index_add_(1, src, torch.einsum("beo,oec->bec", gk, U[:, perm, :]))
        _t(df, (B, Ns, C))[:] = d.float()
    if du:
        D = torch.zeros((Cout, E, C), dtype=torch.float64)
        D[:, perm, :] = torch.einsum("beo,bec->oec", gk, F[:, src, :])
        _t(du, (Cout, E, C))[:] = D.float()


def _edge_mlp_ref(y, dy, x, dx, kc, index, edge_row, E, alpha, W1, b1, Wh0, bh0, Wh1, bh1, n_hidden, Wo, bo, C, grad=False):
    idx, row = _arr(index, (E,), np.int64), _arr(edge_row, (E,), np.int64)
    n_y, n_x = int(idx.max()) + 1, int(row.max()) + 1
    z = torch.cat([_t(y, (n_y, dy))[idx][:, :kc], _t(x, (n_x, dx))[row][:, :kc]], 1).double()
    ps = [_t(W1, (64, 2 * kc)), _t(b1, (64,)), _t(Wh0, (64, 64)), _t(bh0, (64,))]
    if n_hidden == 2:
        ps += [_t(Wh1, (64, 64)), _t(bh1, (64,))]
    ps += [_t(Wo, (C, 64)), _t(bo, (C,))]
    ps = [p.double().requires_grad_(grad) for p in ps]
    h = z
    for i in range(0, len(ps) - 2, 2):
        h = torch.nn.functional.gelu(h @ ps[i].T + ps[i + 1])
    k = h @ ps[-2].T + ps[-1]
    a = _t(alpha, (E,)).double() if alpha else torch.ones(E, dtype=torch.float64)
    return k * a[:, None], ps


def gaot_edge_mlp_fwd(y, dy, x, dx, kc, index, edge_row, E, alpha, W1, b1, Wh0, bh0, Wh1, bh1, n_hidden, Wo, bo, C, w):
    out, _ = _edge_mlp_ref(y, dy, x, dx, kc, index, edge_row, E, alpha, W1, b1, Wh0, bh0, Wh1, bh1, n_hidden, Wo, bo, C)
    _t(w, (E, C))[:] = out.float()


def gaot_edge_mlp_bwd(y, dy, x, dx, kc, index, edge_row, E, alpha, W1, b1, Wh0, bh0, Wh1, bh1, n_hidden, Wo, bo, C, dw, dW1, db1,
                      dWh0, dbh0, dWh1, dbh1, dWo, dbo):
    with torch.enable_grad():
        out, ps = _edge_mlp_ref(y, dy, x, dx, kc, index, edge_row, E, alpha, W1, b1, Wh0, bh0, Wh1, bh1, n_hidden, Wo, bo, C, grad=True)
        out.backward(_t(dw, (E, C)).double())
    dst = [(dW1, (64, 2 * kc)), (db1, (64,)), (dWh0, (64, 64)), (dbh0, (64,))]
    if n_hidden == 2:
        dst += [(dWh1, (64, 64)), (dbh1, (64,))]
    dst += [(dWo, (C, 64)), (dbo, (C,))]
    for (p_, shp), t in zip(dst, ps):
        _t(p_, shp)[:] = t.grad.float()


def gaot_cond_affine_fwd(x, scale, bias, y, B, T, D):
    _arr(y, (B, T, D))[:] = _arr(x, (B, T, D)) * _arr(scale, (B, D))[:, None, :] + _arr(bias, (B, D))[:, None, :]


def gaot_cond_affine_bwd(dy, x, scale, dx, dscale, dbias, B, T, D):
    g, xx = _arr(dy, (B, T, D)).astype(np.float64), _arr(x, (B, T, D)).astype(np.float64)
    _arr(dx, (B, T, D))[:] = (g * _arr(scale, (B, D))[:, None, :]).astype(np.float32)
    _arr(dscale, (B, D))[:] = (g * xx).sum(1).astype(np.float32)
    _arr(dbias, (B, D))[:] = g.sum(1).astype(np.float32)


def gaot_segment_csr(src, row_splits, out, B, E, m, C, reduce):
    rs = _t(row_splits, (m + 1,), np.int64)
    _arr(out, (B, m, C))[:] = orc.segment_csr(_t(src, (B, E, C)), rs, ("sum", "mean", "max")[reduce]).numpy()


def gaot_geo_stats(src, queries, d, index, row_splits, m, feat):
    g, E = _csr(index, row_splits, m)
    idx, rs = g["neighbors_index"], g["neighbors_row_splits"]
    n_s = int(idx.max()) + 1 if E else 0
    S, Q = _t(src, (n_s, d)).double(), _t(queries, (m, d)).double()
    seg, deg = orc._seg_ids(rs)
    n = deg.double()
    nc = n.clamp(min=1)
    pts = S[idx]
    dist = (pts - Q[seg]).norm(dim=1)
    davg = torch.zeros(m, dtype=torch.float64).index_add(0, seg, dist) / nc
    ex2 = torch.zeros(m, dtype=torch.float64).index_add(0, seg, dist ** 2) / nc
    cen = torch.zeros(m, d, dtype=torch.float64).index_add(0, seg, pts) / nc[:, None]
    ctr = pts - cen[seg]
    cov = torch.zeros(m, d, d, dtype=torch.float64).index_add(0, seg, ctr[:, :, None] * ctr[:, None, :]) / nc.view(-1, 1, 1)
    ev = torch.linalg.eigvalsh(cov).flip(dims=[1])
    out = torch.cat([n[:, None], davg[:, None], (ex2 - davg ** 2).clamp(min=0)[:, None], cen - Q, ev], dim=1)
    out[deg == 0] = 0.0
    _arr(feat, (m, 3 + 2 * d))[:] = out.float().numpy()


def gaot_col_standardize(feat, m, F, groups):
    a = _t(feat, (groups, m, F))
    x = a.double()
    mean, std = x.mean(dim=1, keepdim=True), x.std(dim=1, keepdim=True)
    std[std < 1e-6] = 1.0
    a[:] = ((x - mean) / std).float()


# ---- dense --------------------------------------------------------------------------------------
def _strided(ptr, rows, cols, ld):
    return _t(ptr, (rows * ld,))[: (rows - 1) * ld + cols].as_strided((rows, cols), (ld, 1)) if rows and cols else torch.zeros(rows, cols)


def gaot_gemm_f32(A, B, C, M, N, K, lda, ldb, ldc, transA, transB, bias, add, add_rows, ldadd, act, pre, accumulate):
    a = (_strided(A, K, M, lda).T if transA else _strided(A, M, K, lda)).double()
    b = (_strided(B, N, K, ldb).T if transB else _strided(B, K, N, ldb)).double()
    v = a @ b
    if bias:
        v = v + _t(bias, (N,)).double()
    if add:
        v = v + _strided(add, add_rows, N, ldadd).double().repeat(M // add_rows, 1)
    if pre:
        _strided(pre, M, N, ldc)[:] = v.float()
    v = _act(v, act)
    c = _strided(C, M, N, ldc)
    c[:] = (v + c.double()).float() if accumulate else v.float()


def gaot_gemm_bf16(A, B, C, M, N, K, out_bf16, bias, add, add_rows, ldadd, act, pre, accumulate):
    v = _bf(A, (M, K)).double() @ _bf(B, (N, K)).double().T
    if bias:
        v = v + _t(bias, (N,)).double()
    if add:
        v = v + _strided(add, add_rows, N, ldadd).double().repeat(M // add_rows, 1)
    if pre:
        _t(pre, (M, N))[:] = v.float()
    v = _act(v, act)
    if out_bf16:
        _bf(C, (M, N))[:] = v.to(torch.bfloat16)
    else:
        c = _t(C, (M, N))
        c[:] = (v + c.double()).float() if accumulate else v.float()


def gaot_cast_bf16(inp, out, rows, cols, out_t):
    x = _t(inp, (rows, cols))
    if out:
        _bf(out, (rows, cols))[:] = x.to(torch.bfloat16)
    if out_t:
        _bf(out_t, (cols, rows))[:] = x.T.to(torch.bfloat16)


def gaot_transpose_bf16(inp, out_t, rows, cols):
    _bf(out_t, (cols, rows))[:] = _bf(inp, (rows, cols)).T


def gaot_col_sum(x, M, N, out, accumulate):
    o = _arr(out, (N,))
    s = _arr(x, (M, N)).astype(np.float64).sum(0)
    o[:] = (o + s if accumulate else s).astype(np.float32)


def gaot_act_bwd(dy, pre, dx, n, act):
    with torch.enable_grad():
        p = _t(pre, (n,)).double().requires_grad_(True)
        (gr,) = torch.autograd.grad(_act(p, act).sum(), p)
    _arr(dx, (n,))[:] = (_t(dy, (n,)).double() * gr).float().numpy()


def gaot_act_bwd_bf16(dy, dy_is_bf16, pre, dx, n, act):
    with torch.enable_grad():
        p = _t(pre, (n,)).double().requires_grad_(True)
        (gr,) = torch.autograd.grad(_act(p, act).sum(), p)
    g = _bf(dy, (n,)).double() if dy_is_bf16 else _t(dy, (n,)).double()
    _bf(dx, (n,))[:] = (g * gr).to(torch.bfloat16)


def gaot_rmsnorm_fwd(x, gamma, y, y16, rstd, rows, D, eps):
    X = _t(x, (rows, D)).double()
    r = torch.rsqrt(X.pow(2).mean(-1) + eps)
    _arr(rstd, (rows,))[:] = r.float().numpy()
    v = X * r[:, None] * _t(gamma, (D,)).double()
    if y:
        _arr(y, (rows, D))[:] = v.float().numpy()
    if y16:
        _bf(y16, (rows, D))[:] = v.to(torch.bfloat16)


def gaot_rmsnorm_bwd(dy, dy16, x, gamma, rstd, dx, dgamma, rows, D):
    X, gm = _t(x, (rows, D)).double(), _t(gamma, (D,)).double()
    G = torch.zeros(rows, D, dtype=torch.float64)
    if dy:
        G = G + _t(dy, (rows, D)).double()
    if dy16:
        G = G + _bf(dy16, (rows, D)).double()
    r = _t(rstd, (rows,)).double()[:, None]
    dot = (G * gm * X).sum(-1, keepdim=True)
    _arr(dx, (rows, D))[:] = (G * gm * r - X * dot * r ** 3 / D).float().numpy()
    _arr(dgamma, (D,))[:] += (G * X * r).sum(0).float().numpy()


def gaot_swiglu_fwd(ab, u, rows, F):
    a = _t(ab, (rows, 2 * F)).double()
    _arr(u, (rows, F))[:] = (torch.nn.functional.silu(a[:, :F]) * a[:, F:]).float().numpy()


def gaot_swiglu_bwd(du, ab, dab, rows, F):
    with torch.enable_grad():
        a = _t(ab, (rows, 2 * F)).double().requires_grad_(True)
        (torch.nn.functional.silu(a[:, :F]) * a[:, F:]).backward(_t(du, (rows, F)).double())
    _arr(dab, (rows, 2 * F))[:] = a.grad.float().numpy()


def gaot_swiglu_fwd_bf16(ab, u, rows, F):
    a = _bf(ab, (rows, 2 * F)).double()
    _bf(u, (rows, F))[:] = (torch.nn.functional.silu(a[:, :F]) * a[:, F:]).to(torch.bfloat16)


def gaot_swiglu_bwd_bf16(du, ab, dab, rows, F):
    with torch.enable_grad():
        a = _bf(ab, (rows, 2 * F)).double().requires_grad_(True)
        (torch.nn.functional.silu(a[:, :F]) * a[:, F:]).backward(_bf(du, (rows, F)).double())
    _bf(dab, (rows, 2 * F))[:] = a.grad.to(torch.bfloat16)


def _swiglu_blocks(x, F):
    """[rows, 2F] in (a | b) column blocks of 64 -> a [rows, F], b [rows, F] (the layout of gaot_gemm_bf16_swiglu)."""
    v = x.reshape(x.shape[0], F // 32, 2, 32)
    return v[:, :, 0, :].reshape(x.shape[0], F), v[:, :, 1, :].reshape(x.shape[0], F)


def gaot_gemm_bf16_swiglu(A, B, AB, U, M, N, K):
    v = _bf(A, (M, K)).double() @ _bf(B, (N, K)).double().T          # interleaved weight rows -> interleaved columns
    _bf(AB, (M, N))[:] = v.to(torch.bfloat16)
    a, b_ = _swiglu_blocks(v, N // 2)
    _bf(U, (M, N // 2))[:] = (torch.nn.functional.silu(a) * b_).to(torch.bfloat16)


def gaot_gemm_bf16_nn_swiglu_bwd(dY, W, AB, dAB, M, N, K):
    du = _bf(dY, (M, K)).double() @ _bf(W, (K, N)).double()
    a, b_ = _swiglu_blocks(_bf(AB, (M, 2 * N)).double(), N)
    sg = torch.sigmoid(a)
    da = du * b_ * sg * (1 + a * (1 - sg))
    db = du * a * sg
    out = torch.stack([da.reshape(M, N // 32, 32), db.reshape(M, N // 32, 32)], dim=2).reshape(M, 2 * N)
    _bf(dAB, (M, 2 * N))[:] = out.to(torch.bfloat16)


def gaot_gemm_bf16_nn(A, B, C, M, N, K, out_bf16):
    v = _bf(A, (M, K)).double() @ _bf(B, (K, N)).double()
    if out_bf16:
        _bf(C, (M, N))[:] = v.to(torch.bfloat16)
    else:
        _t(C, (M, N))[:] = v.float()


def gaot_gemm_bf16_tn(A, B, C, M, N, R, accumulate):
    v = _bf(A, (R, M)).double().T @ _bf(B, (R, N)).double()
    c = _t(C, (M, N))
    c[:] = (v + c.double()).float() if accumulate else v.float()


def _heads(base, B, T, H, hd, ld):
    return _t(base, (B * T * ld,))[: (B * T - 1) * ld + H * hd].as_strided((B, T, H, hd), (T * ld, ld, hd, 1))


def _attn_ref(q, k, v, H, Hkv, scale):
    rep = H // Hkv
    q, k, v = q.transpose(1, 2), k.transpose(1, 2).repeat_interleave(rep, 1), v.transpose(1, 2).repeat_interleave(rep, 1)
    s = q @ k.transpose(-1, -2) * scale
    return (torch.softmax(s, -1) @ v).transpose(1, 2), torch.logsumexp(s, -1)


def gaot_attention_fwd(q, k, v, o, lse, B, T, H, Hkv, hd, ldq, ldk, scale):
    Q, K, V = _heads(q, B, T, H, hd, ldq).double(), _heads(k, B, T, Hkv, hd, ldk).double(), _heads(v, B, T, Hkv, hd, ldk).double()
    out, l = _attn_ref(Q, K, V, H, Hkv, scale)
    _arr(o, (B, T, H, hd))[:] = out.float().numpy()
    _arr(lse, (B, H, T))[:] = l.float().numpy()


def gaot_attention_bwd(q, k, v, o, lse, dout, delta, dq, dk, dv, B, T, H, Hkv, hd, ldq, ldk, scale):
    with torch.enable_grad():
        Q = _heads(q, B, T, H, hd, ldq).double().requires_grad_(True)
        K = _heads(k, B, T, Hkv, hd, ldk).double().requires_grad_(True)
        V = _heads(v, B, T, Hkv, hd, ldk).double().requires_grad_(True)
        out, _ = _attn_ref(Q, K, V, H, Hkv, scale)
        out.backward(_t(dout, (B, T, H, hd)).double())
    _heads(dq, B, T, H, hd, ldq)[:] = Q.grad.float()
    _heads(dk, B, T, Hkv, hd, ldk)[:] = K.grad.float()
    _heads(dv, B, T, Hkv, hd, ldk)[:] = V.grad.float()


def _bheads(base, B, T, H, hd, ld, col0):
    return _bf(base, (B * T, ld))[:, col0:col0 + H * hd].reshape(B, T, H, hd)


def gaot_attention_fwd_bf16(qkv, o, lse, B, T, H, Hkv, hd, ld, scale):
    Q, K, V = (_bheads(qkv, B, T, H, hd, ld, 0).double(), _bheads(qkv, B, T, Hkv, hd, ld, H * hd).double(),
               _bheads(qkv, B, T, Hkv, hd, ld, (H + Hkv) * hd).double())
    out, l = _attn_ref(Q, K, V, H, Hkv, scale)
    _bf(o, (B, T, H, hd))[:] = out.to(torch.bfloat16)
    _arr(lse, (B, H, T))[:] = l.float().numpy()


def gaot_attention_bwd_bf16(qkv, o, lse, dout, delta, dq_acc, dqkv, B, T, H, Hkv, hd, ld, scale):
    with torch.enable_grad():
        Q = _bheads(qkv, B, T, H, hd, ld, 0).double().requires_grad_(True)
        K = _bheads(qkv, B, T, Hkv, hd, ld, H * hd).double().requires_grad_(True)
        V = _bheads(qkv, B, T, Hkv, hd, ld, (H + Hkv) * hd).double().requires_grad_(True)
        out, _ = _attn_ref(Q, K, V, H, Hkv, scale)
        out.backward(_bf(dout, (B, T, H, hd)).double())
    d = _bf(dqkv, (B * T, ld))
    d[:, :H * hd] = Q.grad.reshape(B * T, H * hd).to(torch.bfloat16)
    d[:, H * hd:(H + Hkv) * hd] = K.grad.reshape(B * T, Hkv * hd).to(torch.bfloat16)
    d[:, (H + Hkv) * hd:(H + 2 * Hkv) * hd] = V.grad.reshape(B * T, Hkv * hd).to(torch.bfloat16)


# ---- split-precision path: planes hold fp16(hi), fp16(lo) of x * scale; products use the three terms the kernel issues ----
def _f16(ptr, shape):
    n = int(np.prod(shape))
    raw = np.ctypeslib.as_array(ctypes.cast(int(ptr), ctypes.POINTER(ctypes.c_uint16)), shape=(n,))
    return torch.from_numpy(raw.view(np.int16)).view(torch.float16).view(shape)


def gaot_split_f16(inp, planes, scale, ws, n):
    x = _t(inp, (n,))
    a = float(x.abs().max())
    sc = 1.0
    if 0.0 < a <= 3.0e38:
        _, e = math.frexp(a)
        sc = math.ldexp(1.0, min(14 - e, 126))
    xs = x * sc                                           # power-of-two scale: exact in fp32
    hi = xs.to(torch.float16)
    lo = (xs - hi.float()).to(torch.float16)
    pl = _f16(planes, (2, n))
    pl[0], pl[1] = hi, lo
    _arr(scale, (2,))[:] = (sc, 1.0 / sc)


def _split3(A, B):
    """(A_lo B_hi + A_hi B_lo + A_hi B_hi) with the contraction over the last axis of A and the first of B"""
    ah, al, bh, bl = A[0].double(), A[1].double(), B[0].double(), B[1].double()
    return al @ bh + ah @ bl + ah @ bh


def gaot_gemm_split(A, B, C, M, N, K, a_mn, b_mn, sa, sb, bias, add, add_rows, ldadd, act, pre, accumulate):
    a = _f16(A, (2, K, M)).transpose(1, 2) if a_mn else _f16(A, (2, M, K))
    b = _f16(B, (2, K, N)) if b_mn else _f16(B, (2, N, K)).transpose(1, 2)
    v = _split3(a, b) * float(_arr(sa, (2,))[1]) * float(_arr(sb, (2,))[1])
    if bias:
        v = v + _t(bias, (N,)).double()
    if add:
        v = v + _strided(add, add_rows, N, ldadd).double().repeat(M // add_rows, 1)
    if pre:
        _t(pre, (M, N))[:] = v.float()
    v = _act(v, act)
    c = _t(C, (M, N))
    c[:] = (v + c.double()).float() if accumulate else v.float()


def _plane_heads(planes, B, T, ld, H, hd, col0):
    p = _f16(planes, (2, B * T, ld))[:, :, col0:col0 + H * hd].double()
    return (p[0] + p[1]).reshape(B, T, H, hd)


def gaot_attention_fwd_split(qkv, qscale, o, lse, B, T, H, Hkv, hd, ld, scale):
    inv = float(_arr(qscale, (2,))[1])
    Q, K, V = (_plane_heads(qkv, B, T, ld, H, hd, 0) * inv, _plane_heads(qkv, B, T, ld, Hkv, hd, H * hd) * inv,
               _plane_heads(qkv, B, T, ld, Hkv, hd, (H + Hkv) * hd) * inv)
    out, l = _attn_ref(Q, K, V, H, Hkv, scale)
    _arr(o, (B, T, H, hd))[:] = out.float().numpy()
    _arr(lse, (B, H, T))[:] = l.float().numpy()


def gaot_attention_bwd_split(qkv, qscale, o, lse, dout, dout_planes, gscale, delta, dqkv, B, T, H, Hkv, hd, ld, scale):
    inv, invg = float(_arr(qscale, (2,))[1]), float(_arr(gscale, (2,))[1])
    with torch.enable_grad():
        Q = (_plane_heads(qkv, B, T, ld, H, hd, 0) * inv).requires_grad_(True)
        K = (_plane_heads(qkv, B, T, ld, Hkv, hd, H * hd) * inv).requires_grad_(True)
        V = (_plane_heads(qkv, B, T, ld, Hkv, hd, (H + Hkv) * hd) * inv).requires_grad_(True)
        out, _ = _attn_ref(Q, K, V, H, Hkv, scale)
        out.backward(_plane_heads(dout_planes, B, T, H * hd, H, hd, 0) * invg)
    d = _t(dqkv, (B * T, ld))
    d[:, :H * hd] = Q.grad.reshape(B * T, H * hd).float()
    d[:, H * hd:(H + Hkv) * hd] = K.grad.reshape(B * T, Hkv * hd).float()
    d[:, (H + Hkv) * hd:(H + 2 * Hkv) * hd] = V.grad.reshape(B * T, Hkv * hd).float()


def gaot_patch_permute(inp, out, B, dims, H, W, D, P, C, to_tokens):
    grid = (H, W) if dims == 2 else (H, W, D)
    nodes, pv = int(np.prod(grid)), P ** dims
    if to_tokens:
        _arr(out, (B, nodes // pv, pv * C))[:] = orc.patchify(_t(inp, (B, nodes, C)), grid, P).numpy()
    else:
        _arr(out, (B, nodes, C))[:] = orc.unpatchify(_t(inp, (B, nodes // pv, pv * C)), grid, P, C).numpy()


# ---- step ---------------------------------------------------------------------------------------
def gaot_mse_fwd_bwd(pred, target, loss, dpred, n, gscale):
    d = _arr(pred, (n,)).astype(np.float64) - _arr(target, (n,)).astype(np.float64)
    _arr(loss, (1,))[:] += np.float32((d * d).mean())
    if dpred:
        _arr(dpred, (n,))[:] = (2.0 * d / n * gscale).astype(np.float32)


def gaot_adamw_step(param, grad, m, v, n, lr, b1, b2, eps, wd, step, gscale):
    p, g = _arr(param, (n,)), _arr(grad, (n,)).astype(np.float64) * gscale
    mm, vv = _arr(m, (n,)), _arr(v, (n,))
    pn = p.astype(np.float64) * (1 - lr * wd)
    m2 = b1 * mm.astype(np.float64) + (1 - b1) * g
    v2 = b2 * vv.astype(np.float64) + (1 - b2) * g * g
    den = np.sqrt(v2) / math.sqrt(1 - b2 ** step) + eps
    p[:] = (pn - lr / (1 - b1 ** step) * m2 / den).astype(np.float32)
    mm[:], vv[:] = m2.astype(np.float32), v2.astype(np.float32)


def gaot_rollout_update(pred, u_in, u_phys, u_next, rows, C, dt, mode, um, us, sm, ss):
    P = _arr(pred, (rows, C)).astype(np.float64)
    m_, s_ = _arr(um, (C,)).astype(np.float64), _arr(us, (C,)).astype(np.float64)
    if mode == 0:
        un = P * s_ + m_
    else:
        ucur = _arr(u_in, (rows, C)) * s_ + m_
        inc = P * _arr(ss, (C,)) + _arr(sm, (C,))
        un = ucur + (dt * inc if mode == 2 else inc)
    _arr(u_phys, (rows, C))[:] = un.astype(np.float32)
    if u_next:
        _arr(u_next, (rows, C))[:] = ((un - m_) / s_).astype(np.float32)


_TABLE = {k: v for k, v in list(globals().items()) if k.startswith("gaot_")}
CALLS = []


def _call(name, *args):
    from gaot_b200 import _lib
    sig = _lib.parse_header()[name][1]
    assert len(args) == len(sig) - 1, f"{name}: {len(args)} arguments for signature {sig} (+stream)"
    for a, kind in zip(args, sig):
        if kind == "f":
            assert isinstance(a, float), (name, a)
        else:
            assert isinstance(a, int) and not isinstance(a, bool), (name, kind, a)
    CALLS.append(name)
    _TABLE[name](*args)


@contextlib.contextmanager
def install():
    """Route gaot_b200's ABI calls to the emulator (CPU tensors) inside the `with` block."""
    from gaot_b200 import _lib, ops
    saved = (_lib.call, ops.call, ops._need_cuda)
    _lib.call, ops.call = _call, _call
    ops._need_cuda = lambda t: None
    CALLS.clear()
    try:
        yield CALLS
    finally:
        _lib.call, ops.call, ops._need_cuda = saved
