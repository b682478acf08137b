"""CPU tests of the host side above the C ABI: the library loads and exports the header's symbols, the
product refuses CPU tensors, and - with the ABI routed to tests/abi_emulator.py - the module wiring,
autograd formulas, vx graph merging and the flat-buffer trainer reproduce the reference's golden vectors."""
import contextlib
import os

import numpy as np
import pytest
import torch

from oracle import cases, gaot_oracle as orc
from tests import abi_emulator
from tests.helpers import (build_model, golden_graphs, golden_params, grad_rel, grad_scale, graphs_to, load_golden, rel_l2)


def test_library_exports_every_declared_symbol():
    from gaot_b200 import _lib
    decl = _lib.parse_header()
    assert len(decl) >= 27
    L = _lib.lib()
    for name in decl:
        assert hasattr(L, name), name
    assert L.gaot_abi_version() == 1


def test_product_path_has_no_cpu_fallback():
    from gaot_b200 import _lib, ops
    with pytest.raises(_lib.GaotLibraryError):
        ops.linear(torch.zeros(4, 4), torch.zeros(4, 4))
    with pytest.raises(_lib.GaotLibraryError):
        ops.radius_csr(torch.zeros(4, 2), torch.zeros(4, 2), 0.1)


@contextlib.contextmanager
def _precision(mode):
    from gaot_b200 import ops
    old = ops.get_precision()
    ops.set_precision(mode)
    try:
        yield
    finally:
        ops.set_precision(old)


def test_nothing_in_the_product_imports_the_oracle():
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gaot_b200")
    for dp, _, fs in os.walk(root):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "abi_emulator" not in src, f


@pytest.mark.parametrize("precision", ["fp32", "ffma"])
@pytest.mark.parametrize("fuse", [True, False])
@pytest.mark.parametrize("name", ["tiny2d", "tiny3d", "plain2d", "vx2d", "dot2d", "npe2d", "pnet2d", "cond2d", "vxmix", "konly2d", "rope2d"])
def test_model_wiring_matches_reference_golden(name, fuse, precision):
    """Both parity modes through the emulated ABI: 'fp32' = split-precision tensor-core entry points (the emulator forms
    the same fp16 hi / lo planes and the same three product terms as the kernels), 'ffma' = plain fp32 entry points."""
    c = cases.CASES[name]
    g = load_golden(name)
    with abi_emulator.install() as calls, _precision(precision):
        from gaot_b200 import ops
        model = build_model(c, "cpu")
        model.encoder.fuse_affine = model.decoder.fuse_affine = fuse
        model.load_state_dict(golden_params(g))
        lat, x, pn, tgt = cases.case_inputs(name)
        enc = dec = None
        if c.get("vx"):
            enc, dec = golden_graphs(g, "enc", True), golden_graphs(g, "dec", True)
        out = model(lat, x, pn, encoder_nbrs=enc, decoder_nbrs=dec, condition=cases.case_condition(name))
        loss = ops.mse_loss(out, tgt)
        loss.backward()
        assert ("gaot_attention_bwd_split" if precision == "fp32" else "gaot_attention_bwd") in calls
        assert ("gaot_gemm_split" in calls) == (precision == "fp32")
        fusable = fuse and c["magno"]["lifting_channels"] % 4 == 0        # tiny3d has C = 6: falls back
        assert ("gaot_lift_aggregate_bwd" in calls) == fusable
        if not fusable:
            assert "gaot_agno_aggregate" in calls and "gaot_agno_weight_grad" in calls
        elif len(c["magno"].get("scales", [1.0])) == 1:
            assert "gaot_agno_dot_bwd" in calls and "gaot_agno_dot_fwd" in calls       # B < 16: row walk
    assert rel_l2(out, g["out"]) < 1e-5
    assert abs(float(loss) - float(g["loss64"])) < 1e-5 * abs(float(g["loss64"]))
    gs = grad_scale(g, g["keys"])
    for k, p in model.named_parameters():
        assert grad_rel(p.grad, g["grad." + k], gs) < 1e-4, k
    if not c.get("vx"):   # graphs the model built itself == the reference's
        for tag, mod in (("enc", model.encoder), ("dec", model.decoder)):
            mine = mod.neighbor_cache["graph"][1]
            for a, b in zip(mine, golden_graphs(g, tag, False)):
                assert torch.equal(a["neighbors_index"], b["neighbors_index"])
                assert torch.equal(a["neighbors_row_splits"], b["neighbors_row_splits"])


def test_vx_without_precomputed_graphs_searches_per_sample():
    c = dict(cases.CASES["vx2d"])
    c["magno"] = dict(c["magno"], precompute_edges=False)
    g = load_golden("vx2d")
    with abi_emulator.install():
        model = build_model(c, "cpu")
        model.load_state_dict(golden_params(g))
        lat, x, pn, _ = cases.case_inputs("vx2d")
        out = model(lat, x, pn)
    assert rel_l2(out, g["out"]) < 1e-5


def test_error_conventions_match_reference():
    c = cases.CASES["tiny2d"]
    with abi_emulator.install():
        model = build_model(c, "cpu")
        lat, x, pn, _ = cases.case_inputs("tiny2d")
        with pytest.raises(ValueError):
            model(lat, x[:, :1], pn)                      # wrong coord_dim
        with pytest.raises(ValueError):
            model(lat, x, pn[:, :-1])                     # node count mismatch
        with pytest.raises(ValueError):
            model.process(torch.zeros(2, 63, 16))         # token count != H*W
    from gaot_b200.model import MAGNOConfig
    with pytest.raises(ValueError):
        MAGNOConfig(coord_dim=4)
    with pytest.raises(ValueError):
        MAGNOConfig(sampling_strategy="ratio")


def _rand_csr(m, n_src, max_deg, seed):
    g = torch.Generator().manual_seed(seed)
    deg = torch.randint(0, max_deg + 1, (m,), generator=g)
    rs = torch.zeros(m + 1, dtype=torch.int64)
    rs[1:] = torch.cumsum(deg, 0)
    # ascending distinct sources inside a row (what the radius search produces)
    idx = torch.cat([torch.sort(torch.randperm(n_src, generator=g)[:int(d)]).values for d in deg]) if int(rs[-1]) else \
        torch.zeros(0, dtype=torch.int64)
    return {"neighbors_index": idx, "neighbors_row_splits": rs}


def _check_sub_csr(out, src):
    """`out` is a valid CSR whose every row is an order-preserving subset of the same row of `src`."""
    rs, ro = src["neighbors_row_splits"], out["neighbors_row_splits"]
    assert ro.shape == rs.shape and int(ro[0]) == 0 and int(ro[-1]) == out["neighbors_index"].numel()
    assert bool((ro[1:] >= ro[:-1]).all())
    for r in range(rs.numel() - 1):
        full = src["neighbors_index"][rs[r]:rs[r + 1]].tolist()
        sub = out["neighbors_index"][ro[r]:ro[r + 1]].tolist()
        it = iter(full)
        assert all(any(a == b for b in it) for a in sub), r       # subsequence test


@pytest.mark.parametrize("tag,use_attn", [("attn", True), ("mean", False)])
def test_agno_kernel_only_and_integration_weights_match_golden(tag, use_attn):
    """AGNO options outside MAGNO's call sites (agno.py:150-273) through the emulated ABI: f_y=None (kernel-only transform)
    and the f_y transform against the reference's own AGNO module; `weights=` against the oracle restatement (the
    reference's line only broadcasts for C == E; tests/golden/agno.npz comes from `oracle/make_golden.py agno`)."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "agno.npz"))
    y, x, f_y, wts, layers, _ = cases.agno_case()
    nbrs = {"neighbors_index": torch.from_numpy(g["index"]), "neighbors_row_splits": torch.from_numpy(g["splits"])}
    with abi_emulator.install():
        from gaot_b200.model.layers.agno import AGNO
        m = AGNO(channel_mlp_layers=layers, use_attn=use_attn, attention_type="cosine", coord_dim=2)
        m.load_state_dict({k: torch.from_numpy(g[f"{tag}.param.{k}"]) for k in m.state_dict()})
        assert rel_l2(m(y, nbrs, x, f_y=None), g[f"{tag}.kernel_only"]) < 1e-5
        assert rel_l2(m(y, nbrs, x, f_y=f_y), g[f"{tag}.full"]) < 1e-5
        assert rel_l2(m(y, nbrs, x, f_y=f_y[0]), g[f"{tag}.full"][0]) < 1e-5                     # un-batched f_y [n, C]
        assert rel_l2(m(y, nbrs, x, f_y=None, weights=wts), g[f"{tag}.kernel_only_weighted"]) < 1e-5
        out = m(y, nbrs, x, f_y=f_y, weights=wts)
        assert rel_l2(out, g[f"{tag}.full_weighted"]) < 1e-5
        out.sum().backward()                                                                     # the option paths are differentiable
        assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters())
        with pytest.raises(AssertionError):
            m(y, nbrs, x, f_y=f_y, weights=wts[:, None])                                         # agno.py:254
        with pytest.raises(NotImplementedError):
            AGNO(channel_mlp_layers=layers, transform_type="nonlinear")                          # shape-broken in the reference
        AGNO(channel_mlp_layers=layers, transform_type="linear_kernelonly")                      # == 'linear' (agno.py:229-247)


def test_edge_drop_identity_cases_follow_the_reference():
    """edge_drop.py:38-42, 55-56, 71-72, 79-80, 98-99: every early return hands back the SAME dict."""
    from gaot_b200.model.layers.utils.edge_drop import apply_edge_drop_csr as drop
    g = _rand_csr(50, 40, 9, 0)
    assert drop(g, None) is g
    assert drop(g, "ratio", sample_ratio=0.5, training=False) is g
    assert drop(g, "ratio", sample_ratio=None) is g
    assert drop(g, "ratio", sample_ratio=1.0) is g
    assert drop(g, "max_neighbors", max_neighbors=None) is g
    assert drop(g, "max_neighbors", max_neighbors=9) is g            # no row above the cap
    assert drop(g, "something_else") is g
    empty = {"neighbors_index": torch.zeros(0, dtype=torch.int64), "neighbors_row_splits": torch.zeros(6, dtype=torch.int64)}
    assert drop(empty, "ratio", sample_ratio=0.5) is empty
    bad = {"neighbors_index": g["neighbors_index"], "neighbors_row_splits": g["neighbors_row_splits"].flip(0)}
    with pytest.raises(ValueError):
        drop(bad, "ratio", sample_ratio=0.5)
    with pytest.raises(ValueError):
        drop(bad, "max_neighbors", max_neighbors=2)


def test_edge_drop_ratio_sampling_statistics():
    """edge_drop.py:50-68: i.i.d. Bernoulli(sample_ratio) per edge; result is a valid sub-CSR."""
    from gaot_b200.model.layers.utils.edge_drop import apply_edge_drop_csr as drop
    g = _rand_csr(400, 300, 40, 1)
    E = g["neighbors_index"].numel()
    torch.manual_seed(0)
    out = drop(g, "ratio", sample_ratio=0.3)
    _check_sub_csr(out, g)
    kept = out["neighbors_index"].numel()
    assert abs(kept - 0.3 * E) < 5 * (0.3 * 0.7 * E) ** 0.5           # 5 sigma of Binomial(E, 0.3)
    assert out["neighbors_row_splits"].dtype == g["neighbors_row_splits"].dtype


def test_edge_drop_max_neighbors_cap_and_uniformity():
    """edge_drop.py:70-97: rows above the cap keep EXACTLY max_neighbors edges, a uniformly random subset; other rows are
    untouched.  Uniformity: over many draws every edge of an over-full row is kept with frequency cap/deg - also for a
    row index large enough that a fp32 `rand + 2*row` key would lose its random bits (the round-1 bug)."""
    from gaot_b200.model.layers.utils.edge_drop import apply_edge_drop_csr as drop
    g = _rand_csr(300, 200, 30, 2)
    torch.manual_seed(1)
    out = drop(g, "max_neighbors", max_neighbors=7)
    _check_sub_csr(out, g)
    deg = g["neighbors_row_splits"][1:] - g["neighbors_row_splits"][:-1]
    deg_o = out["neighbors_row_splits"][1:] - out["neighbors_row_splits"][:-1]
    assert torch.equal(deg_o, torch.clamp(deg, max=7))
    rs, ro = g["neighbors_row_splits"], out["neighbors_row_splits"]
    for r in torch.where(deg <= 7)[0].tolist():
        assert torch.equal(out["neighbors_index"][ro[r]:ro[r + 1]], g["neighbors_index"][rs[r]:rs[r + 1]])
    # one over-full row placed at a very large row index: 12 edges, cap 4 -> each edge kept with probability 1/3
    m = 5_000_000
    rs_big = torch.zeros(m + 1, dtype=torch.int64)
    rs_big[m:] = 12
    big = {"neighbors_index": torch.arange(12), "neighbors_row_splits": rs_big}
    hits = torch.zeros(12)
    trials = 150
    torch.manual_seed(2)
    for _ in range(trials):
        o = drop(big, "max_neighbors", max_neighbors=4)
        assert o["neighbors_index"].numel() == 4
        hits[o["neighbors_index"]] += 1
    p = 1.0 / 3.0
    assert float((hits / trials - p).abs().max()) < 5 * (p * (1 - p) / trials) ** 0.5
    assert len({tuple(drop(big, "max_neighbors", max_neighbors=4)["neighbors_index"].tolist()) for _ in range(20)}) > 5


def test_edge_drop_inside_the_encoder_training_vs_eval():
    """MAGNO applies the sampling in training mode only (magno.py:279, 309): eval output is deterministic and equals the
    no-sampling model; training output differs (edges were dropped)."""
    c = dict(cases.CASES["tiny2d"])
    c["magno"] = dict(c["magno"], sampling_strategy="ratio", sample_ratio=0.5)
    g = load_golden("tiny2d")
    with abi_emulator.install():
        model = build_model(c, "cpu")
        model.load_state_dict(golden_params(g))
        lat, x, pn, _ = cases.case_inputs("tiny2d")
        model.eval()
        with torch.no_grad():
            e1 = model(lat, x, pn)
        assert rel_l2(e1, g["out"]) < 1e-5                          # eval: sampling off -> the reference golden
        model.train()
        torch.manual_seed(3)
        with torch.no_grad():
            t1 = model(lat, x, pn)
        assert rel_l2(t1, g["out"]) > 1e-3


@pytest.mark.parametrize("n_hidden,kc,C,attn", [(2, 2, 64, True), (1, 3, 48, False)])
def test_fused_edge_mlp_wiring_equals_unfused(n_hidden, kc, C, attn):
    """AGNO.edge_weights in bf16 mode takes the fused edge-weight entry points (emulated in fp64 here) when the kernel MLP has
    the instantiated shape; values and parameter gradients equal the unfused chain of the fp32 mode."""
    from gaot_b200 import ops
    from gaot_b200.model.layers.agno import AGNO
    data, q, r = cases.nbr_case("rand2d" if kc == 2 else "rand3d")
    g = orc.radius_csr(data, q, r)
    with abi_emulator.install() as calls:
        torch.manual_seed(0)
        agno = AGNO(channel_mlp_layers=[2 * kc] + [64] * (n_hidden + 1) + [C], use_attn=attn, attention_type="cosine", coord_dim=kc)
        plan = ops.GraphPlan(g["neighbors_index"], g["neighbors_row_splits"], data.shape[0])
        gw = cases.normal((plan.E, C), 5)
        res = {}
        for mode in ("bf16", "ffma"):
            with _precision(mode):
                agno.zero_grad()
                w = agno.edge_weights(data, q, plan)
                w.backward(gw)
                res[mode] = (w.detach().clone(), [p.grad.clone() for p in agno.parameters()])
        assert "gaot_edge_mlp_fwd" in calls and "gaot_edge_mlp_bwd" in calls
    assert rel_l2(res["bf16"][0], res["ffma"][0]) < 1e-5
    for a, b in zip(res["bf16"][1], res["ffma"][1]):
        assert rel_l2(a, b) < 1e-5
    assert ops.edge_mlp_supported(4, 64, 2, 64) and not ops.edge_mlp_supported(4, 16, 2, 16) and not ops.edge_mlp_supported(4, 64, 3, 64)
    from gaot_b200 import _lib
    for args in ((4, 64, 2, 64), (6, 64, 1, 48), (4, 16, 2, 16), (5, 64, 2, 64), (4, 64, 2, 24)):
        assert bool(_lib.lib().gaot_edge_mlp_supported(*args)) == ops.edge_mlp_supported(*args)


def test_train_step_flat_buffers_and_adamw():
    name = "tiny2d"
    c = cases.CASES[name]
    g = load_golden(name)
    with abi_emulator.install():
        from gaot_b200.trainer import TrainStep
        model = build_model(c, "cpu")
        model.load_state_dict(golden_params(g))
        lat, x, pn, tgt = cases.case_inputs(name)
        step = TrainStep(model, lr=1e-3, weight_decay=1e-2, distributed=False)
        losses = [float(step.train_step((pn, tgt), lat, x)) for _ in range(2)]
    cfg = orc.make_config(c["latent"], c["magno"], c["transformer"], c["attn"])
    p = {k: v.double() for k, v in golden_params(g).items()}
    m = {k: torch.zeros_like(v) for k, v in p.items()}
    v_ = {k: torch.zeros_like(v) for k, v in p.items()}
    for it in range(1, 3):
        pr = {k: t.clone().requires_grad_(True) for k, t in p.items()}
        lo = orc.train_step_loss(pr, cfg, lat.double(), x.double(), pn.double(), tgt.double())
        lo.backward()
        assert abs(losses[it - 1] - float(lo)) < 1e-5 * abs(float(lo))
        p, m, v_ = orc.adamw_step({k: t.detach() for k, t in pr.items()}, {k: t.grad for k, t in pr.items()}, m, v_, it, 1e-3)
    for k, t in model.state_dict().items():
        assert rel_l2(t, p[k]) < 1e-5, k


def _dp_worker(rank, world, port, q, row_shard=False):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    name = "tiny2d"
    c = cases.CASES[name]
    g = load_golden(name)
    with abi_emulator.install():
        from gaot_b200 import ops
        from gaot_b200.trainer import TrainStep
        if row_shard:                                     # MAGNO rows sharded over the ranks (ops.enable_edge_sharding)
            ops.SHARD_EDGES_MIN = 1
            ops.enable_edge_sharding(True)
        model = build_model(c, "cpu")
        sd = golden_params(g)
        if rank != 0:                                     # replicas must be re-synchronised from rank 0
            sd = {k: v + 1.0 for k, v in sd.items()}
        model.load_state_dict(sd)
        lat, x, pn, tgt = cases.case_inputs(name)
        per = c["B"] // world
        sl = slice(rank * per, (rank + 1) * per)
        step = TrainStep(model, lr=1e-3, weight_decay=1e-2, bucket_bytes=16 << 10)   # small buckets: several reductions in flight
        assert len(step._buckets) > 3 and sum(hi - lo for lo, hi, _, _ in step._buckets) == len(step.params)
        assert step._buckets[0][3] == step.flat_grad.numel() and step._buckets[-1][2] == 0           # cover the flat buffer
        assert all(a[2] == b[3] for a, b in zip(step._buckets, step._buckets[1:]))                   # ... contiguously
        loss = step.fixed_coords_step(lat, x, pn[sl].contiguous(), tgt[sl].contiguous())
        assert all(step._fired)
    if rank == 0:
        q.put(({k: v.detach().numpy().copy() for k, v in model.state_dict().items()}, float(loss)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("row_shard", [False, True])
def test_data_parallel_step_equals_single_process_gloo(row_shard):
    """world_size 2 over gloo: sample-sharded step (sum all-reduce, 1/world in AdamW) == full-batch step; with row_shard the
    MAGNO operators additionally run row-sharded (every rank owns half of the latent / query rows for both ranks' samples:
    all-gather in, all-to-all out, inverse exchanges in the backward, partial parameter gradients summed by the all-reduce)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + 7 * int(row_shard)) % 2000
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q, row_shard)) for r in range(2)]
    for p in procs:
        p.start()
    sd_dp, _ = q.get(timeout=300)
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    name = "tiny2d"
    c = cases.CASES[name]
    g = load_golden(name)
    with abi_emulator.install():
        from gaot_b200.trainer import TrainStep
        model = build_model(c, "cpu")
        model.load_state_dict(golden_params(g))
        lat, x, pn, tgt = cases.case_inputs(name)
        TrainStep(model, lr=1e-3, weight_decay=1e-2, distributed=False).fixed_coords_step(lat, x, pn, tgt)
    for k, t in model.state_dict().items():
        assert rel_l2(torch.from_numpy(sd_dp[k]), t) < 1e-5, k


@pytest.mark.parametrize("mode", ["time_der", "residual", "output"])
def test_rollout_matches_reference_golden(mode):
    """GAOT.autoregressive_predict (gaot.py:307-476) through the ABI emulator vs the reference's own fp64 output
    (tests/golden/rollout.npz): de-normalised predictions are returned, [u | c | ...] input layout with stats['c']."""
    c, x0, tv, stats = cases.rollout_case()
    g = load_golden("rollout")
    with abi_emulator.install():
        model = build_model(c, "cpu").eval()
        model.load_state_dict(golden_params(g))
        lat, x, _, _ = cases.case_inputs(cases.ROLLOUT["base"])
        xb = torch.cat([x0, torch.zeros(x0.shape[0], x0.shape[1], 2)], -1)      # trailing dataset time channels are ignored
        got = model.autoregressive_predict(xb, cases.ROLLOUT["time_indices"], tv, stats, mode, latent_tokens_coord=lat,
                                           fixed_coord=x)
    want = torch.from_numpy(g["pred." + mode])
    assert got.shape == want.shape
    for i in range(want.shape[1]):
        assert rel_l2(got[:, i], want[:, i]) < 1e-5, (mode, i)


def test_rollout_rejects_unknown_stepper_mode():
    c, x0, tv, stats = cases.rollout_case()
    with abi_emulator.install():
        model = build_model(c, "cpu").eval()
        with pytest.raises(ValueError):
            model.autoregressive_predict(x0, [0, 1], tv, stats, "euler")


def test_rollout_with_conditional_norm_feeds_time_as_condition():
    """gaot.py:392-408: with use_conditional_norm the time-difference channel is dropped from the encoder input and the
    conditioning value is x_input[..., 0, -2:-1]."""
    c = dict(cases.CASES["cond2d"], cin=4, cout=3)       # encoder sees u(3) + t0; dt is held back
    with abi_emulator.install():
        torch.manual_seed(0)
        model = build_model(c, "cpu").eval()
        lat, x, _, _ = cases.case_inputs("cond2d")
        u0 = cases.normal((3, 300, 3), 41)
        stats = {"u": {"mean": torch.tensor([0.1, 0.2, 0.3]), "std": torch.tensor([1.1, 0.9, 1.2])},
                 "der": {"mean": torch.tensor([0.0, 0.1, -0.1]), "std": torch.tensor([0.5, 0.6, 0.7])},
                 "start_time": {"mean": 0.5, "std": 0.3}, "time_diffs": {"mean": 0.1, "std": 0.05}}
        tv = torch.linspace(0, 1, 21)
        got = model.autoregressive_predict(u0, [0, 3, 5], tv, stats, "time_der", latent_tokens_coord=lat, fixed_coord=x,
                                           use_conditional_norm=True)
        cur = u0
        for i, (a, b) in enumerate(((0, 3), (3, 5))):
            t0, dt = float(tv[a]), float(tv[b] - tv[a])
            t0n = (t0 - 0.5) / 0.3
            inp = torch.cat([cur, torch.full((3, 300, 1), t0n)], -1)
            pred = model(lat, x, inp, condition=torch.full((3, 1), t0n))
            phys, cur = orc.rollout_update(pred, cur, dt, stats, "time_der")
            assert rel_l2(got[:, i], phys) < 1e-6


def test_bf16_mode_wiring_close_to_reference():
    """precision='bf16' routes the processor's Linear layers through the tensor-core ABI (emulated here);
    bf16 operand rounding keeps the output within a few 1e-3 of the fp64 reference on the tiny model."""
    name = "tiny2d"
    c = cases.CASES[name]
    g = load_golden(name)
    with abi_emulator.install() as calls:
        from gaot_b200 import ops
        ops.set_precision("bf16")
        try:
            model = build_model(c, "cpu")
            model.load_state_dict(golden_params(g))
            lat, x, pn, tgt = cases.case_inputs(name)
            out = model(lat, x, pn)
            loss = ops.mse_loss(out, tgt)
            loss.backward()
        finally:
            ops.set_precision("fp32")
        assert "gaot_gemm_bf16" in calls and "gaot_gemm_bf16_tn" in calls and "gaot_attention_bwd_bf16" in calls
        # the gated FFN (hidden 128 = 4 blocks of 32) runs with the SwiGLU inside the GEMM epilogues, forward and backward
        assert "gaot_gemm_bf16_swiglu" in calls and "gaot_gemm_bf16_nn_swiglu_bwd" in calls and "gaot_swiglu_bwd_bf16" not in calls
    assert rel_l2(out, g["out"]) < 2e-2
    for k, p in model.named_parameters():
        assert rel_l2(p.grad, g["grad." + k]) < 6e-2, k


def test_prepared_weights_are_used_once_and_never_outlive_a_forward():
    """ops.prepare_weight registers the bf16 copy of a weight for the next tensor-core linear on that very tensor object
    (GAOT.forward prepares the processor's weights on a side stream); clear_prepared() drops every entry."""
    with abi_emulator.install() as calls:
        from gaot_b200 import ops
        ops.set_precision("bf16")
        try:
            x, W = cases.normal((64, 32), 51), torch.nn.Parameter(cases.normal((16, 32), 52))
            y_ref = ops.linear(x, W, tc=True)
            n_cast = calls.count("gaot_cast_bf16")
            ops.prepare_weight(W)
            assert calls.count("gaot_cast_bf16") == n_cast + 1
            y = ops.linear(x, W, tc=True)                          # casts x only: the weight's bf16 copy is the prepared one
            assert calls.count("gaot_cast_bf16") == n_cast + 2
            assert torch.equal(y, y_ref)
            y.sum().backward()                                     # the prepared copy also serves dX in the backward
            assert W.grad is not None
            ops.clear_prepared()
            n_now = calls.count("gaot_cast_bf16")
            ops.linear(x, W, tc=True)
            assert calls.count("gaot_cast_bf16") == n_now + 2      # x and W again
        finally:
            ops.clear_prepared()
            ops.set_precision("fp32")
