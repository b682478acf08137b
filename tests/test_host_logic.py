"""CPU tests of the host side above the C ABI: the library loads and exports the header's symbols, the
product refuses CPU tensors, and - with the ABI routed to tests/abi_emulator.py - the module wiring,
autograd formulas, vx graph merging and the flat-buffer trainer reproduce the reference's golden vectors."""
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


def test_nothing_in_the_product_imports_the_oracle():
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gaot_b200")
    for dp, _, fs in os.walk(root):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "abi_emulator" not in src, f


@pytest.mark.parametrize("fuse", [True, False])
@pytest.mark.parametrize("name", ["tiny2d", "tiny3d", "plain2d", "vx2d", "dot2d", "npe2d", "pnet2d", "cond2d", "vxmix"])
def test_model_wiring_matches_reference_golden(name, fuse):
    c = cases.CASES[name]
    g = load_golden(name)
    with abi_emulator.install() as calls:
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
        assert "gaot_attention_bwd" in calls
        fusable = fuse and c["magno"]["lifting_channels"] % 4 == 0        # tiny3d has C = 6: falls back
        assert ("gaot_lift_aggregate_bwd" in calls) == fusable
        if not fusable:
            assert "gaot_agno_aggregate" in calls and "gaot_agno_weight_grad" in calls
        elif len(c["magno"].get("scales", [1.0])) == 1:
            assert "gaot_agno_dot_bwd" in calls
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


def _dp_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    name = "tiny2d"
    c = cases.CASES[name]
    g = load_golden(name)
    with abi_emulator.install():
        from gaot_b200.trainer import TrainStep
        model = build_model(c, "cpu")
        sd = golden_params(g)
        if rank != 0:                                     # replicas must be re-synchronised from rank 0
            sd = {k: v + 1.0 for k, v in sd.items()}
        model.load_state_dict(sd)
        lat, x, pn, tgt = cases.case_inputs(name)
        per = c["B"] // world
        sl = slice(rank * per, (rank + 1) * per)
        step = TrainStep(model, lr=1e-3, weight_decay=1e-2)
        loss = step.fixed_coords_step(lat, x, pn[sl].contiguous(), tgt[sl].contiguous())
    if rank == 0:
        q.put(({k: v.detach().numpy().copy() for k, v in model.state_dict().items()}, float(loss)))
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_step_equals_single_process_gloo():
    """world_size 2 over gloo: sample-sharded step (sum all-reduce, 1/world in AdamW) == full-batch step."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
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
def test_rollout_matches_oracle_update_rules(mode):
    c = dict(cases.CASES["tiny2d"], cin=5, cout=3)       # u(3) + t0 + dt
    with abi_emulator.install():
        torch.manual_seed(0)
        model = build_model(c, "cpu").eval()
        lat, x, _, _ = cases.case_inputs("tiny2d")
        u0 = cases.normal((2, 300, 3), 31)
        stats = {"u": {"mean": torch.tensor([0.1, 0.2, 0.3]), "std": torch.tensor([1.1, 0.9, 1.2])},
                 "der": {"mean": torch.tensor([0.0, 0.1, -0.1]), "std": torch.tensor([0.5, 0.6, 0.7])},
                 "res": {"mean": torch.tensor([0.05, -0.1, 0.2]), "std": torch.tensor([0.3, 0.4, 0.2])},
                 "start_time": {"mean": 0.5, "std": 0.3}, "time_diffs": {"mean": 0.1, "std": 0.05}}
        tv = torch.linspace(0, 1, 21)
        got = model.autoregressive_predict(u0, [0, 2, 4], tv, stats, mode, latent_tokens_coord=lat, fixed_coord=x)
        # oracle: same forward through the emulator, update rule from oracle.rollout_update
        cur = u0
        for i, (a, b) in enumerate(((0, 2), (2, 4))):
            t0, dt = float(tv[a]), float(tv[b] - tv[a])
            inp = torch.cat([cur, torch.full((2, 300, 1), (t0 - 0.5) / 0.3), torch.full((2, 300, 1), (dt - 0.1) / 0.05)], -1)
            pred = model(lat, x, inp)
            cur = orc.rollout_update(pred, cur, dt, stats, mode)
            assert rel_l2(got[:, i], cur) < 1e-6


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
            cur = orc.rollout_update(pred, cur, dt, stats, "time_der")
            assert rel_l2(got[:, i], cur) < 1e-6


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
        assert "gaot_swiglu_bwd_bf16" in calls
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
