"""Pins the CPU oracle (oracle/gaot_oracle.py) to golden vectors produced by the unmodified reference
(oracle/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import cases, gaot_oracle as orc
from tests.helpers import golden_graphs, golden_params, golden_shapes, grad_rel, grad_scale, load_golden, rel_l2

FULL = ["tiny2d", "tiny3d", "vx2d", "plain2d", "dot2d", "npe2d", "pnet2d", "cond2d", "vxmix", "konly2d", "rope2d"]


def _cfg(name):
    c = cases.CASES[name]
    return c, orc.make_config(c["latent"], c["magno"], c["transformer"], c["attn"])


@pytest.mark.parametrize("name", cases.NBR_CASES)
def test_radius_csr_matches_reference_native(name):
    g = load_golden("neighbors")
    data, q, r = cases.nbr_case(name)
    o = orc.radius_csr(data, q, r)
    assert np.array_equal(o["neighbors_index"].numpy(), g[f"{name}.index"].astype(np.int64))
    assert np.array_equal(o["neighbors_row_splits"].numpy(), g[f"{name}.splits"].astype(np.int64))


def test_radius_csr_torch_cluster_semantics():
    # strict '<' drops exact-boundary neighbours; cap keeps the first k ascending
    data, q, r = cases.nbr_case("lattice2d")
    inc = orc.radius_csr(data, q, r)
    strict = orc.radius_csr(data, q, r, inclusive=False)
    cap = orc.radius_csr(data, q, r, max_num_neighbors=3)
    assert strict["neighbors_index"].numel() < inc["neighbors_index"].numel()
    rs, rc = inc["neighbors_row_splits"], cap["neighbors_row_splits"]
    for m in range(q.shape[0]):
        full = inc["neighbors_index"][rs[m]:rs[m + 1]]
        assert torch.equal(cap["neighbors_index"][rc[m]:rc[m + 1]], full[:3])


@pytest.mark.parametrize("name", FULL)
def test_forward_backward_matches_reference_fp64(name):
    c, cfg = _cfg(name)
    g = load_golden(name)
    lat, x, pn, tgt = [t.double() for t in cases.case_inputs(name)]
    p = {k: v.double().requires_grad_(True) for k, v in golden_params(g).items()}
    enc = golden_graphs(g, "enc", bool(c.get("vx")))
    dec = golden_graphs(g, "dec", bool(c.get("vx")))
    cond = cases.case_condition(name)
    out = orc.gaot_forward(p, cfg, lat, x, pn, None, enc, dec, cond.double() if cond is not None else None)
    assert rel_l2(out, g["out"]) < 2e-7          # fixtures are stored as fp32
    loss = torch.nn.functional.mse_loss(out, tgt)
    assert abs(float(loss.detach()) - float(g["loss64"])) < 1e-12 * max(1.0, abs(float(g["loss64"])))
    loss.backward()
    gs = grad_scale(g, g["keys"])
    for k in p:
        assert grad_rel(p[k].grad, g["grad." + k], gs) < 2e-7, k


@pytest.mark.parametrize("name", ["tiny2d", "tiny3d", "plain2d"])
def test_graphs_rebuilt_by_oracle_equal_reference(name):
    c, cfg = _cfg(name)
    g = load_golden(name)
    lat, x, _, _ = cases.case_inputs(name)
    mc = cfg.args.magno
    for tag, (dd, qq) in {"enc": (x, lat), "dec": (lat, x)}.items():
        ref = golden_graphs(g, tag, False)
        mine = orc.build_graphs(mc, dd, qq)
        for a, b in zip(mine, ref):
            assert torch.equal(a["neighbors_index"], b["neighbors_index"])
            assert torch.equal(a["neighbors_row_splits"], b["neighbors_row_splits"])


def test_cfg2_summary_matches_reference():
    """BASELINE cfg-2 geometry (N=9216, latent 64x64, r=0.033): params regenerated from the seed."""
    c, cfg = _cfg("cfg2")
    g = load_golden("cfg2")
    lat, x, pn, tgt = cases.case_inputs("cfg2")
    p = {k: v.double() for k, v in cases.synth_params(golden_shapes(g), 7 + c["seed"]).items()}
    enc = orc.build_graphs(cfg.args.magno, x, lat)
    assert int(enc[0]["neighbors_index"].numel()) == int(g["enc_E"])
    assert int(enc[0]["neighbors_index"].sum()) == int(g["enc_index_sum"])
    assert np.array_equal(enc[0]["neighbors_row_splits"][::64].numpy(), g["enc_splits_sub"])
    dec = orc.build_graphs(cfg.args.magno, lat, x)
    assert int(dec[0]["neighbors_index"].sum()) == int(g["dec_index_sum"])
    with torch.no_grad():
        out = orc.gaot_forward(p, cfg, lat.double(), x.double(), pn.double(), None, enc, dec)
    assert rel_l2(out[:, ::37, :], g["out_sub"]) < 2e-7
    assert abs(float(out.norm()) - float(g["out_norm"])) < 1e-9 * float(g["out_norm"])


def test_csr_transpose_roundtrip():
    data, q, r = cases.nbr_case("rand2d")
    g = orc.radius_csr(data, q, r)
    t = orc.csr_transpose(g, data.shape[0])
    rs = g["neighbors_row_splits"]
    row = torch.repeat_interleave(torch.arange(q.shape[0]), rs[1:] - rs[:-1])
    ts = t["neighbors_row_splits"]
    src = torch.repeat_interleave(torch.arange(data.shape[0]), ts[1:] - ts[:-1])
    assert torch.equal(g["neighbors_index"][t["edge_perm"]], src)
    assert torch.equal(row[t["edge_perm"]], t["neighbors_index"])


def test_adamw_matches_torch():
    torch.manual_seed(0)
    w = torch.randn(7, 5, dtype=torch.float64)
    ref = torch.nn.Parameter(w.clone())
    opt = torch.optim.AdamW([ref], lr=1e-3, weight_decay=1e-2)
    params, m, v = {"w": w.clone()}, {"w": torch.zeros_like(w)}, {"w": torch.zeros_like(w)}
    for step in range(1, 4):
        gr = torch.randn(7, 5, dtype=torch.float64)
        ref.grad = gr.clone()
        opt.step()
        params, m, v = orc.adamw_step(params, {"w": gr}, m, v, step, 1e-3)
        assert rel_l2(params["w"], ref.detach()) < 1e-14


@pytest.mark.parametrize("mode", ["output", "residual", "time_der"])
def test_rollout_matches_reference_golden(mode):
    """oracle.autoregressive_predict vs the reference's GAOT.autoregressive_predict output (de-normalised predictions,
    [u | c | ...] input layout with stats['c'], gaot.py:307-476)."""
    c, x0, tv, stats = cases.rollout_case()
    g = load_golden("rollout")
    cfg = orc.make_config(c["latent"], c["magno"], c["transformer"], c["attn"])
    p = {k: v.double() for k, v in golden_params(g).items()}
    st64 = {k: {kk: (vv.double() if torch.is_tensor(vv) else vv) for kk, vv in v.items()} for k, v in stats.items()}
    lat, x, _, _ = cases.case_inputs(cases.ROLLOUT["base"])
    got = orc.autoregressive_predict(p, cfg, x0.double(), cases.ROLLOUT["time_indices"], tv.double(), st64, mode, lat.double(),
                                     x.double())
    assert rel_l2(got, g["pred." + mode]) < 2e-7
