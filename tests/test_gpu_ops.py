"""GPU parity of the individual C-ABI operators against the CPU oracle / fp64 torch (run with -m gpu).
Integer outputs are compared bit-exact; floating point within the tolerance written in each test."""
import numpy as np
import pytest
import torch

from oracle import cases, gaot_oracle as orc
from tests.helpers import load_golden, rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _ops():
    from gaot_b200 import ops
    return ops


@pytest.mark.parametrize("name", cases.NBR_CASES)
def test_radius_csr_bit_exact_vs_reference_golden(name):
    ops = _ops()
    g = load_golden("neighbors")
    data, q, r = cases.nbr_case(name)
    out = ops.radius_csr(data.to(DEV), q.to(DEV), r)
    assert out["neighbors_index"].dtype == torch.int64 and out["neighbors_row_splits"].dtype == torch.int64
    assert np.array_equal(out["neighbors_index"].cpu().numpy(), g[f"{name}.index"].astype(np.int64))
    assert np.array_equal(out["neighbors_row_splits"].cpu().numpy(), g[f"{name}.splits"].astype(np.int64))


@pytest.mark.parametrize("name", cases.NBR_CASES)
def test_radius_csr_grid_bit_exact_vs_reference_golden(name):
    """The cell-binned search against the reference `native` goldens: exact-boundary lattices, duplicates, empty rows,
    queries far outside the data box."""
    ops = _ops()
    g = load_golden("neighbors")
    data, q, r = cases.nbr_case(name)
    out = ops.radius_csr(data.to(DEV), q.to(DEV), r, method="grid")
    assert np.array_equal(out["neighbors_index"].cpu().numpy(), g[f"{name}.index"].astype(np.int64))
    assert np.array_equal(out["neighbors_row_splits"].cpu().numpy(), g[f"{name}.splits"].astype(np.int64))
    strict = ops.radius_csr(data.to(DEV), q.to(DEV), r, inclusive=False, method="grid")
    want = orc.radius_csr(data, q, r, inclusive=False)
    assert torch.equal(strict["neighbors_index"].cpu(), want["neighbors_index"])
    with pytest.raises(ValueError):
        ops.radius_csr(data.to(DEV), q.to(DEV), r, max_num_neighbors=4, method="grid")


@pytest.mark.parametrize("d,n,m,r", [(2, 30000, 7000, 0.02), (3, 40000, 5000, 0.09), (2, 5000, 20000, 0.05), (3, 300, 300, 1.5)])
def test_radius_csr_grid_equals_sweep(d, n, m, r):
    """Same bit-exact graph from both searches on larger random clouds (uniform data, clustered queries; a radius larger
    than the box in the last case), incl. the batched forms the vx-mode graph builder uses."""
    ops = _ops()
    data = cases.uniform_points(n, d, 91).to(DEV)
    q = cases.clustered_points(m, d, 92).to(DEV)
    a, b = ops.radius_csr(data, q, r, method="sweep"), ops.radius_csr(data, q, r, method="grid")
    assert torch.equal(a["neighbors_row_splits"], b["neighbors_row_splits"]) and torch.equal(a["neighbors_index"], b["neighbors_index"])
    assert int(a["neighbors_index"].numel()) > 0
    xs = torch.stack([cases.clustered_points(min(n, 4000), d, 93 + s) for s in range(3)]).to(DEV)
    lat = cases.uniform_points(900, d, 97).to(DEV)
    for dd, qq in ((xs, lat), (lat, xs)):
        sa, sb = ops.radius_csr(dd, qq, r, method="sweep"), ops.radius_csr(dd, qq, r, method="grid")
        for ga, gb in zip(sa, sb):
            assert torch.equal(ga["neighbors_row_splits"], gb["neighbors_row_splits"]) and torch.equal(ga["neighbors_index"], gb["neighbors_index"])


@pytest.mark.parametrize("name", ["lattice2d", "rand3d"])
def test_radius_csr_torch_cluster_semantics(name):
    ops = _ops()
    data, q, r = cases.nbr_case(name)
    for incl, cap in [(False, None), (True, 3), (False, 32)]:
        want = orc.radius_csr(data, q, r, inclusive=incl, max_num_neighbors=cap)
        got = ops.radius_csr(data.to(DEV), q.to(DEV), r, inclusive=incl, max_num_neighbors=cap)
        assert torch.equal(got["neighbors_index"].cpu(), want["neighbors_index"])
        assert torch.equal(got["neighbors_row_splits"].cpu(), want["neighbors_row_splits"])


def test_radius_csr_batched_equals_per_sample():
    ops = _ops()
    lat = cases.latent_grid((8, 8))
    xs = torch.stack([cases.clustered_points(500, 2, 40 + b) for b in range(3)])
    enc = ops.radius_csr(xs.to(DEV), lat.to(DEV), 0.3)          # data batched
    dec = ops.radius_csr(lat.to(DEV), xs.to(DEV), 0.3)          # queries batched
    for b in range(3):
        for got, want in ((enc[b], orc.radius_csr(xs[b], lat, 0.3)), (dec[b], orc.radius_csr(lat, xs[b], 0.3))):
            assert torch.equal(got["neighbors_index"].cpu(), want["neighbors_index"])
            assert torch.equal(got["neighbors_row_splits"].cpu(), want["neighbors_row_splits"])


def test_radius_csr_large_property():
    """BASELINE cfg-2 size: the decoder graph is the transpose of the encoder graph, rows ascending."""
    ops = _ops()
    x = cases.uniform_points(9216, 2, 0).to(DEV)
    lat = cases.latent_grid((64, 64)).to(DEV)
    enc = ops.radius_csr(x, lat, 0.033)
    dec = ops.radius_csr(lat, x, 0.033)
    g = load_golden("cfg2")
    assert int(enc["neighbors_index"].numel()) == int(g["enc_E"])
    assert int(enc["neighbors_index"].sum()) == int(g["enc_index_sum"])
    assert int(dec["neighbors_index"].sum()) == int(g["dec_index_sum"])
    assert np.array_equal(enc["neighbors_row_splits"][::64].cpu().numpy(), g["enc_splits_sub"])
    assert np.array_equal(dec["neighbors_row_splits"][::128].cpu().numpy(), g["dec_splits_sub"])
    plan = ops.GraphPlan(enc["neighbors_index"], enc["neighbors_row_splits"], 9216)
    ts, tr, tp = plan.transposed()
    assert torch.equal(ts, dec["neighbors_row_splits"]) and torch.equal(tr, dec["neighbors_index"])


def test_csr_transpose_and_edge_rows_vs_oracle():
    ops = _ops()
    data, q, r = cases.nbr_case("rand2d")
    g = orc.radius_csr(data, q, r)
    t = orc.csr_transpose(g, data.shape[0])
    plan = ops.GraphPlan(g["neighbors_index"].to(DEV), g["neighbors_row_splits"].to(DEV), data.shape[0])
    ts, tr, tp = plan.transposed()
    assert torch.equal(ts.cpu(), t["neighbors_row_splits"])
    assert torch.equal(tr.cpu(), t["neighbors_index"])
    assert torch.equal(tp.cpu(), t["edge_perm"])
    rs = g["neighbors_row_splits"]
    row = torch.repeat_interleave(torch.arange(q.shape[0]), rs[1:] - rs[:-1])
    assert torch.equal(plan.edge_row.cpu(), row)


@pytest.mark.parametrize("M,N,K", [(300, 64, 4), (257, 130, 67), (1000, 1, 64), (64, 256, 1000), (5, 7, 3), (2048, 512, 256),
                                   (5000, 64, 64), (4100, 48, 6), (3000, 7, 64), (31245, 64, 4)])
def test_linear_fwd_bwd_vs_fp64(M, N, K):
    ops = _ops()
    gen = torch.Generator().manual_seed(M * 7 + N)
    x = torch.randn(M, K, generator=gen)
    W = torch.randn(N, K, generator=gen) / K ** 0.5
    b = torch.randn(N, generator=gen)
    R = 5 if M % 5 == 0 else M
    add = torch.randn(R, N, generator=gen)
    for act, fact in ((ops.ACT_NONE, lambda t: t), (ops.ACT_GELU, torch.nn.functional.gelu), (ops.ACT_RELU, torch.relu)):
        xs = [t.clone().to(DEV).requires_grad_(True) for t in (x, W, b, add)]
        y = ops.linear(xs[0], xs[1], xs[2], add=xs[3], act=act)
        xr = [t.clone().double().requires_grad_(True) for t in (x, W, b, add)]
        yr = fact(xr[0] @ xr[1].T + xr[2] + xr[3].repeat(M // R, 1))
        assert rel_l2(y, yr) < 2e-6
        gy = torch.randn(M, N, generator=gen)
        y.backward(gy.to(DEV))
        yr.backward(gy.double())
        for a, r_ in zip(xs, xr):
            assert rel_l2(a.grad, r_.grad) < 5e-6


def test_aggregate_fwd_bwd_vs_oracle():
    ops = _ops()
    data, q, r = cases.nbr_case("rand2d")
    g = orc.radius_csr(data, q, r)
    E = g["neighbors_index"].numel()
    for C, B in ((64, 5), (16, 2), (48, 9), (6, 3), (128, 1)):
        f = cases.normal((B, data.shape[0], C), 5)
        w = cases.normal((E, C), 6)
        fr, wr = f.double().requires_grad_(True), w.double().requires_grad_(True)
        want = orc.segment_csr(wr.unsqueeze(0) * fr[:, g["neighbors_index"], :], g["neighbors_row_splits"], "sum")
        fg, wg = f.to(DEV).requires_grad_(True), w.to(DEV).requires_grad_(True)
        plan = ops.GraphPlan(g["neighbors_index"].to(DEV), g["neighbors_row_splits"].to(DEV), data.shape[0])
        got = ops.agno_aggregate(fg, wg, plan)
        assert rel_l2(got, want) < 1e-6
        go = cases.normal(tuple(want.shape), 7)
        want.backward(go.double())
        got.backward(go.to(DEV))
        assert rel_l2(fg.grad, fr.grad) < 1e-6
        assert rel_l2(wg.grad, wr.grad) < 2e-6


def test_segment_csr_dropin():
    ops = _ops()
    data, q, r = cases.nbr_case("lattice2d")          # has empty rows
    g = orc.radius_csr(data, q, r)
    E = g["neighbors_index"].numel()
    src = cases.normal((3, E, 5), 8)
    for red in ("sum", "mean", "max"):
        want = orc.segment_csr(src, g["neighbors_row_splits"], red)
        got = ops.segment_csr(src.to(DEV), g["neighbors_row_splits"].to(DEV), red)
        assert rel_l2(got, want) < 1e-6
    want = orc.segment_csr(src[0, :, 0], g["neighbors_row_splits"], "max")
    got = ops.segment_csr(src[0, :, 0].contiguous().to(DEV), g["neighbors_row_splits"].to(DEV), "max")
    assert rel_l2(got, want) < 1e-6


@pytest.mark.parametrize("name", ["rand2d", "rand3d", "lattice2d"])
def test_edge_softmax_and_geo_stats_vs_oracle(name):
    ops = _ops()
    data, q, r = cases.nbr_case(name)
    g = orc.radius_csr(data, q, r)
    plan = ops.GraphPlan(g["neighbors_index"].to(DEV), g["neighbors_row_splits"].to(DEV), data.shape[0])
    seg, _ = orc._seg_ids(g["neighbors_row_splits"])
    qd, yd = q.double()[seg], data.double()[g["neighbors_index"]]
    s = (torch.nn.functional.normalize(qd, dim=-1) * torch.nn.functional.normalize(yd, dim=-1)).sum(-1)
    ex = torch.exp(s - orc.segment_csr(s, g["neighbors_row_splits"], "max")[seg])
    want = ex / orc.segment_csr(ex, g["neighbors_row_splits"], "sum")[seg]
    got = ops.edge_cosine_softmax(data.to(DEV), q.to(DEV), data.shape[1], plan)
    assert rel_l2(got, want) < 2e-6
    z = ops.edge_inputs(data.to(DEV), q.to(DEV), plan)
    assert torch.equal(z.cpu(), torch.cat([data[g["neighbors_index"]], q[seg]], dim=-1))
    feat = ops.geo_stats(data.to(DEV), q.to(DEV), plan)
    want = orc.geo_stat_features(data.double(), q.double(), g)
    # the smallest covariance eigenvalue of near-degenerate neighbourhoods is the noisiest column
    assert rel_l2(feat, want) < 5e-6


@pytest.mark.parametrize("B,T,H,Hkv,hd", [(2, 100, 2, 2, 32), (1, 257, 4, 2, 16), (2, 64, 2, 1, 64)])
def test_attention_fwd_bwd_vs_oracle(B, T, H, Hkv, hd):
    ops = _ops()
    W = (H + 2 * Hkv) * hd
    qkv = cases.normal((B * T, W), 9)
    go = cases.normal((B * T, H * hd), 10)

    def ref(qkv64):
        q = qkv64[:, :H * hd].view(B, T, H, hd).transpose(1, 2)
        k = qkv64[:, H * hd:(H + Hkv) * hd].view(B, T, Hkv, hd).transpose(1, 2).repeat_interleave(H // Hkv, dim=1)
        v = qkv64[:, (H + Hkv) * hd:].view(B, T, Hkv, hd).transpose(1, 2).repeat_interleave(H // Hkv, dim=1)
        p = torch.softmax(q @ k.transpose(-1, -2) / hd ** 0.5, dim=-1)
        return (p @ v).transpose(1, 2).reshape(B * T, H * hd)

    a = qkv.double().requires_grad_(True)
    want = ref(a)
    want.backward(go.double())
    b = qkv.to(DEV).requires_grad_(True)
    got = ops.attention(b, B, T, H, Hkv, hd)
    got.backward(go.to(DEV))
    assert rel_l2(got, want) < 2e-6
    assert rel_l2(b.grad, a.grad) < 5e-6


def test_rmsnorm_swiglu_patch_mse_adamw():
    ops = _ops()
    x = cases.normal((77, 96), 11)
    gm = 1 + 0.1 * cases.normal((96,), 12)
    go = cases.normal((77, 96), 13)
    a, ag = x.double().requires_grad_(True), gm.double().requires_grad_(True)
    want = a * torch.rsqrt(a.pow(2).mean(-1, keepdim=True) + 1e-6) * ag
    want.backward(go.double())
    b, bg = x.to(DEV).requires_grad_(True), gm.to(DEV).requires_grad_(True)
    got = ops.rmsnorm(b, bg, 1e-6)
    got.backward(go.to(DEV))
    y32, y16 = ops.rmsnorm(b.detach(), bg.detach(), 1e-6, "both")
    assert torch.equal(y32, got.detach()) and torch.equal(y16, got.detach().to(torch.bfloat16))
    assert rel_l2(got, want) < 1e-6 and rel_l2(b.grad, a.grad) < 2e-6 and rel_l2(bg.grad, ag.grad) < 2e-6
    # swiglu
    ab = cases.normal((50, 64), 14)
    a = ab.double().requires_grad_(True)
    want = torch.nn.functional.silu(a[:, :32]) * a[:, 32:]
    want.backward(go[:50, :32].double())
    b = ab.to(DEV).requires_grad_(True)
    got = ops.swiglu(b)
    got.backward(go[:50, :32].contiguous().to(DEV))
    assert rel_l2(got, want) < 1e-6 and rel_l2(b.grad, a.grad) < 2e-6
    # patchify round trip and parity, 2-D and 3-D
    for grid, P, C in (((8, 4), 2, 3), ((4, 4, 6), 2, 5)):
        n = int(np.prod(grid))
        t = cases.normal((2, n, C), 15)
        want = orc.patchify(t, grid, P)
        got = ops.patchify(t.to(DEV), grid, P)
        assert torch.equal(got.cpu(), want)
        assert torch.equal(ops.unpatchify(got, grid, P, C).cpu(), t)
    # mse
    p, y = cases.normal((3, 50, 2), 16), cases.normal((3, 50, 2), 17)
    a = p.double().requires_grad_(True)
    want = torch.nn.functional.mse_loss(a, y.double())
    want.backward()
    b = p.to(DEV).requires_grad_(True)
    got = ops.mse_loss(b, y.to(DEV))
    got.backward()
    assert abs(float(got) - float(want)) < 1e-6 * float(want) and rel_l2(b.grad, a.grad) < 1e-6
    # adamw vs torch.optim.AdamW (the oracle's adamw_step is pinned to it in test_oracle_golden.py)
    w = cases.normal((1000,), 18)
    ref = torch.nn.Parameter(w.clone().double())
    opt = torch.optim.AdamW([ref], lr=1e-3, weight_decay=1e-2)
    pw, m, v = w.to(DEV), torch.zeros(1000, device=DEV), torch.zeros(1000, device=DEV)
    for step in range(1, 4):
        gr = cases.normal((1000,), 19 + step)
        ref.grad = gr.double()
        opt.step()
        ops.adamw_step(pw, gr.to(DEV), m, v, step, 1e-3)
    assert rel_l2(pw, ref.detach()) < 1e-6


def test_fused_lift_aggregate_and_dot_vs_oracle():
    ops = _ops()
    data, q, r = cases.nbr_case("rand2d")
    g = orc.radius_csr(data, q, r)
    E, n, m = g["neighbors_index"].numel(), data.shape[0], q.shape[0]
    plan = ops.GraphPlan(g["neighbors_index"].to(DEV), g["neighbors_row_splits"].to(DEV), n)
    # B a multiple of 8 takes the sample-minor gather (pn_t) for every Cin; the others the per-sample gather
    for C, B, Cin in ((64, 11, 1), (48, 3, 3), (16, 8, 4), (64, 1, 2), (64, 32, 1), (64, 16, 2), (32, 24, 3)):
        pn, Wl, bl, w = cases.normal((B, n, Cin), 31), cases.normal((C, Cin), 32), cases.normal((C,), 33), cases.normal((E, C), 34)
        ref = [t.double().requires_grad_(True) for t in (Wl, bl, w)]
        f = pn.double() @ ref[0].T + ref[1]
        want = orc.segment_csr(ref[2].unsqueeze(0) * f[:, g["neighbors_index"], :], g["neighbors_row_splits"], "sum")
        mine = [t.to(DEV).requires_grad_(True) for t in (Wl, bl, w)]
        got = ops.lift_aggregate(pn.to(DEV), mine[0], mine[1], mine[2], plan)
        assert rel_l2(got, want) < 2e-6
        go = cases.normal(tuple(want.shape), 35)
        want.backward(go.double())
        got.backward(go.to(DEV))
        for a, b in zip(mine, ref):
            assert rel_l2(a.grad, b.grad) < 5e-6


@pytest.mark.parametrize("row_walk", [False, True])
@pytest.mark.parametrize("name", ["rand2d", "dec2d", "lattice2d", "rand3d", "empty"])
def test_agno_dot_both_walks_vs_oracle(name, row_walk):
    """Decoder aggregate + affine read-out, forward and backward, in both traversal orders: by source node on the transposed
    graph (default: gaot_agno_dot_{fwd,bwd}_t) and by destination row (gaot_agno_dot_{fwd,bwd}); graphs with empty rows,
    sources without edges, duplicate points and degrees above the 8 edges a pass holds in registers."""
    ops = _ops()
    data, q, r = cases.nbr_case(name)
    g = orc.radius_csr(data, q, r)
    E, n, m = g["neighbors_index"].numel(), data.shape[0], q.shape[0]
    plan = ops.GraphPlan(g["neighbors_index"].to(DEV), g["neighbors_row_splits"].to(DEV), n)
    old = ops.ROW_WALK
    ops.ROW_WALK = row_walk
    try:
        for C, B, Cout in ((64, 11, 1), (48, 3, 3), (16, 8, 2), (64, 32, 1), (32, 1, 4), (48, 16, 2), (16, 40, 3), (32, 33, 4)):
            f, u = cases.normal((B, n, C), 36), cases.normal((Cout, E, C), 37)
            fr, ur = f.double().requires_grad_(True), u.double().requires_grad_(True)
            want = orc.segment_csr(torch.einsum("oec,bec->beo", ur, fr[:, g["neighbors_index"], :]), g["neighbors_row_splits"], "sum")
            fg, ug = f.to(DEV).requires_grad_(True), u.to(DEV).requires_grad_(True)
            got = ops.agno_dot(fg, ug, plan)
            go = cases.normal(tuple(want.shape), 38)
            want.backward(go.double())
            got.backward(go.to(DEV))
            if E == 0:
                assert float(got.abs().max()) == 0.0 and float(fg.grad.abs().max()) == 0.0
                continue
            assert rel_l2(got, want) < 2e-6
            assert rel_l2(fg.grad, fr.grad) < 2e-6 and rel_l2(ug.grad, ur.grad) < 5e-6
    finally:
        ops.ROW_WALK = old


@pytest.mark.parametrize("name", ["rand2d", "dec2d", "lattice2d", "empty"])
def test_edge_dot_softmax_fwd_bwd_vs_fp64(name):
    """Dot-product attention weights (agno.py:214-217) and their gradients against an fp64 autograd restatement,
    incl. empty rows and the row scale's gradient (gaot_row_dot)."""
    ops = _ops()
    data, q, r = cases.nbr_case(name)
    g = orc.radius_csr(data, q, r)
    idx, rs = g["neighbors_index"], g["neighbors_row_splits"]
    m, n, E, A, C = q.shape[0], data.shape[0], int(idx.numel()), 64, 16
    plan = ops.GraphPlan(idx.to(DEV), rs.to(DEV), n)
    gen = torch.Generator().manual_seed(5)
    Q, K = torch.randn(m, A, generator=gen), torch.randn(n, A, generator=gen)
    kap, gout = torch.randn(E, C, generator=gen), torch.randn(E, C, generator=gen)
    scale = 1.0 / A ** 0.5
    qd, kd, kp = [t.clone().to(DEV).requires_grad_(True) for t in (Q, K, kap)]
    if E == 0:                                            # no edge at all: the forward must be a no-op that returns an empty vector
        assert ops.edge_dot_softmax(qd, kd, plan, scale).numel() == 0
        return
    w = ops.scale_rows(kp, ops.edge_dot_softmax(qd, kd, plan, scale))
    w.backward(gout.to(DEV))
    qr, kr, kpr = [t.clone().double().requires_grad_(True) for t in (Q, K, kap)]
    seg = torch.repeat_interleave(torch.arange(m), rs[1:] - rs[:-1])
    s = (qr[seg] * kr[idx]).sum(-1) * scale
    ex = torch.exp(s - orc.segment_csr(s, rs, "max")[seg])
    alpha = ex / orc.segment_csr(ex, rs, "sum")[seg]
    wr = kpr * alpha[:, None]
    wr.backward(gout.double())
    assert rel_l2(w, wr) < 2e-6
    assert rel_l2(kp.grad, kpr.grad) < 2e-6
    assert rel_l2(qd.grad, qr.grad) < 2e-5
    assert rel_l2(kd.grad, kr.grad) < 2e-5


@pytest.mark.parametrize("B,T,D", [(3, 16, 64), (2, 257, 40), (32, 1024, 256)])
def test_cond_affine_fwd_bwd_vs_fp64(B, T, D):
    """Time-conditioned affine correction (mlp.py:121-124): y = x * s[b] + o[b], gradients incl. the per-sample sums."""
    ops = _ops()
    gen = torch.Generator().manual_seed(B * 31 + D)
    x, s, o, g = torch.randn(B, T, D, generator=gen), torch.randn(B, D, generator=gen), torch.randn(B, D, generator=gen), torch.randn(B, T, D, generator=gen)
    a = [t.clone().to(DEV).requires_grad_(True) for t in (x, s, o)]
    y = ops.cond_affine(*a)
    y.backward(g.to(DEV))
    r = [t.clone().double().requires_grad_(True) for t in (x, s, o)]
    yr = r[0] * r[1][:, None, :] + r[2][:, None, :]
    yr.backward(g.double())
    assert rel_l2(y, yr) < 1e-6
    for got, want in zip(a, r):
        assert rel_l2(got.grad, want.grad) < 5e-6


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("tag,use_attn", [("attn", True), ("mean", False)])
def test_agno_kernel_only_and_integration_weights_vs_golden(tag, use_attn, precision):
    """AGNO options outside MAGNO's call sites (agno.py:150-273): f_y=None (kernel-only transform) and the f_y transform
    against the reference's own AGNO module, `weights=` against the oracle restatement (tests/golden/agno.npz).
    Tolerance: 1e-5 in the parity mode; 2e-2 in 'bf16' mode, where the kernel MLP's 16 -> 8 layer takes bf16 operands."""
    import os
    ops = _ops()
    from gaot_b200.model.layers.agno import AGNO
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "agno.npz"))
    y, x, f_y, wts, layers, _ = [t.to(DEV) if torch.is_tensor(t) else t for t in cases.agno_case()]
    nbrs = {"neighbors_index": torch.from_numpy(g["index"]).to(DEV), "neighbors_row_splits": torch.from_numpy(g["splits"]).to(DEV)}
    ops.set_precision(precision)
    tol = 1e-5 if precision == "fp32" else 2e-2
    try:
        m = AGNO(channel_mlp_layers=layers, use_attn=use_attn, attention_type="cosine", coord_dim=2).to(DEV)
        m.load_state_dict({k: torch.from_numpy(g[f"{tag}.param.{k}"]).to(DEV) for k in m.state_dict()})
        assert rel_l2(m(y, nbrs, x, f_y=None), g[f"{tag}.kernel_only"]) < tol
        assert rel_l2(m(y, nbrs, x, f_y=f_y), g[f"{tag}.full"]) < tol
        assert rel_l2(m(y, nbrs, x, f_y=None, weights=wts), g[f"{tag}.kernel_only_weighted"]) < tol
        out = m(y, nbrs, x, f_y=f_y, weights=wts)
        assert rel_l2(out, g[f"{tag}.full_weighted"]) < tol
        out.sum().backward()
        assert all(p.grad is not None and bool(torch.isfinite(p.grad).all()) for p in m.parameters())
    finally:
        ops.set_precision("fp32")
