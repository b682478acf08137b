"""GPU parity of the whole GAOT forward/backward (gaot_b200.model, through the C ABI) against golden
vectors produced by the unmodified reference (tests/golden, oracle/make_golden.py).  Run with -m gpu.

Tolerances: forward <= 1e-5 relative L2 (BASELINE.json north_star).  Gradients: per tensor
max(1e-4, 10 x the fp32 noise floor of the reference arithmetic for that tensor), the floor being the
`gerr32` array stored with each fixture (oracle fp32 vs fp64 on the same graph)."""
import numpy as np
import pytest
import torch

from oracle import cases, gaot_oracle as orc
from tests.helpers import (build_model, golden_graphs, golden_params, golden_shapes, grad_rel, grad_scale, graphs_to, load_golden,
                           rel_l2)

pytestmark = pytest.mark.gpu
DEV = "cuda"
FWD_TOL = 1e-5


@pytest.fixture(params=["fp32", "ffma"])
def parity_mode(request):
    """Both parity modes of the processor: 'fp32' = split-precision tcgen05 kernels (three fp16 terms per product),
    'ffma' = fp32 FFMA kernels.  Same tolerances."""
    from gaot_b200 import ops
    ops.set_precision(request.param)
    yield request.param
    ops.set_precision("fp32")


def _run(name, use_golden_graphs, fuse=True):
    c = cases.CASES[name]
    g = load_golden(name)
    model = build_model(c)
    model.encoder.fuse_affine = model.decoder.fuse_affine = fuse
    model.load_state_dict({k: v.to(DEV) for k, v in golden_params(g).items()})
    model.train()
    lat, x, pn, tgt = [t.to(DEV) for t in cases.case_inputs(name)]
    enc = dec = None
    if use_golden_graphs:
        vx = bool(c.get("vx"))
        enc, dec = graphs_to(golden_graphs(g, "enc", vx), DEV), graphs_to(golden_graphs(g, "dec", vx), DEV)
        model.encoder.precompute_edges = model.decoder.precompute_edges = True
    inter = {}
    hooks = [model.encoder.register_forward_hook(lambda m, i, o: inter.__setitem__("enc_out", o.detach())),
             model.processor.register_forward_hook(lambda m, i, o: inter.__setitem__("proc_out", o.detach()))]
    cond = cases.case_condition(name)
    out = model(lat, x, pn, encoder_nbrs=enc, decoder_nbrs=dec, condition=cond.to(DEV) if cond is not None else None)
    for h in hooks:
        h.remove()
    from gaot_b200 import ops
    loss = ops.mse_loss(out, tgt)
    loss.backward()
    return c, g, model, out, loss, inter


@pytest.mark.parametrize("fuse", [True, False])
@pytest.mark.parametrize("name,golden_graph", [("tiny2d", False), ("tiny2d", True), ("tiny3d", False),
                                               ("plain2d", False), ("vx2d", True), ("dot2d", False), ("npe2d", False), ("pnet2d", False),
                                               ("cond2d", False), ("vxmix", True), ("konly2d", False), ("rope2d", False)])
def test_forward_backward_matches_reference(name, golden_graph, fuse, parity_mode):
    c, g, model, out, loss, inter = _run(name, golden_graph, fuse)
    assert rel_l2(inter["enc_out"], g["enc_out"]) < FWD_TOL
    assert rel_l2(out, g["out"]) < FWD_TOL
    assert abs(float(loss) - float(g["loss64"])) < 1e-5 * abs(float(g["loss64"]))
    floors = {str(k): float(e) for k, e in zip(g["keys"], g["gerr32"])}
    worst = 0.0
    gs = grad_scale(g, g["keys"])
    for k, p in model.named_parameters():
        assert p.grad is not None, k
        err = grad_rel(p.grad, g["grad." + k], gs)
        tol = max(1e-4, 10 * floors[k])
        assert err < tol, (k, err, tol)
        worst = max(worst, err)
    print(f"{name}: out {rel_l2(out, g['out']):.2e} worst grad {worst:.2e}")


def test_vx_graphs_built_on_gpu_match_reference():
    from gaot_b200 import ops
    c = cases.CASES["vx2d"]
    g = load_golden("vx2d")
    lat, x, _, _ = cases.case_inputs("vx2d")
    r = c["magno"]["radius"]
    enc = ops.radius_csr(x.to(DEV), lat.to(DEV), r)
    dec = ops.radius_csr(lat.to(DEV), x.to(DEV), r)
    for b in range(c["B"]):
        for mine, ref in ((enc[b], golden_graphs(g, "enc", True)[b][0]), (dec[b], golden_graphs(g, "dec", True)[b][0])):
            assert torch.equal(mine["neighbors_index"].cpu(), ref["neighbors_index"])
            assert torch.equal(mine["neighbors_row_splits"].cpu(), ref["neighbors_row_splits"])


def test_cfg2_matches_reference_summary(parity_mode):
    """BASELINE cfg-2 geometry (N=9216, latent 64x64, r=0.033, 3.4 M parameters), B=2."""
    c = cases.CASES["cfg2"]
    g = load_golden("cfg2")
    model = build_model(c)
    params = cases.synth_params(golden_shapes(g), 7 + c["seed"])
    model.load_state_dict({k: v.to(DEV) for k, v in params.items()})
    lat, x, pn, tgt = [t.to(DEV) for t in cases.case_inputs("cfg2")]
    inter = {}
    model.encoder.register_forward_hook(lambda m, i, o: inter.__setitem__("enc", o.detach()))
    model.processor.register_forward_hook(lambda m, i, o: inter.__setitem__("proc", o.detach()))
    out = model(lat, x, pn)
    assert rel_l2(inter["enc"][:, ::41, ::5], g["enc_sub"]) < FWD_TOL
    assert rel_l2(out[:, ::37, :], g["out_sub"]) < FWD_TOL
    assert abs(float(out.double().norm()) - float(g["out_norm"])) < 1e-5 * float(g["out_norm"])
    from gaot_b200 import ops
    loss = ops.mse_loss(out, tgt)
    loss.backward()
    assert abs(float(loss) - float(g["loss64"])) < 1e-5 * float(g["loss64"])
    floors = {str(k): float(e) for k, e in zip(g["keys"], g["gerr32"])}
    norms = {str(k): float(n) for k, n in zip(g["keys"], g["grad_norms"])}
    grads = dict(model.named_parameters())
    for k in norms:
        got = float(grads[k].grad.double().norm())
        assert abs(got - norms[k]) < max(1e-4, 10 * floors[k]) * norms[k], (k, got, norms[k])
    for k in [kk[5:] for kk in g if kk.startswith("grad.")]:
        assert rel_l2(grads[k].grad, g["grad." + k]) < max(1e-4, 10 * floors[k]), k


def test_train_steps_match_oracle_adamw():
    """3 optimizer steps of the restated trainer loop (optimizers.py:252-255) vs the oracle in fp64."""
    from gaot_b200.trainer import TrainStep
    name = "tiny2d"
    c = cases.CASES[name]
    g = load_golden(name)
    model = build_model(c)
    model.load_state_dict({k: v.to(DEV) for k, v in golden_params(g).items()})
    lat, x, pn, tgt = cases.case_inputs(name)
    step = TrainStep(model, lr=1e-3, weight_decay=1e-2)
    cfg = orc.make_config(c["latent"], c["magno"], c["transformer"], c["attn"])
    p = {k: v.double() for k, v in golden_params(g).items()}
    m = {k: torch.zeros_like(v) for k, v in p.items()}
    v_ = {k: torch.zeros_like(v) for k, v in p.items()}
    enc, dec = orc.build_graphs(cfg.args.magno, x, lat), orc.build_graphs(cfg.args.magno, lat, x)
    for it in range(1, 4):
        loss = step.fixed_coords_step(lat.to(DEV), x.to(DEV), pn.to(DEV), tgt.to(DEV))
        pr = {k: t.clone().requires_grad_(True) for k, t in p.items()}
        lo = orc.train_step_loss(pr, cfg, lat.double(), x.double(), pn.double(), tgt.double(), enc, dec)
        lo.backward()
        p, m, v_ = orc.adamw_step({k: t.detach() for k, t in pr.items()}, {k: t.grad for k, t in pr.items()}, m, v_, it, 1e-3)
        assert abs(float(loss) - float(lo)) < 2e-5 * abs(float(lo)), it
    sd = model.state_dict()
    for k in p:
        assert rel_l2(sd[k], p[k]) < 2e-5, k


@pytest.mark.parametrize("use_graph", [False, True])
def test_autoregressive_rollout_matches_oracle(use_graph):
    """10-step no-grad rollout (gaot.py:307-476, stepper 'time_der') vs the oracle's restatement in fp64; the return value
    is the de-normalised prediction of every step."""
    c = dict(cases.CASES["tiny2d"], cin=5, cout=3)       # u(3) + t0 + dt
    torch.manual_seed(0)
    model = build_model(c).eval()
    lat, x, _, _ = cases.case_inputs("tiny2d")
    u0 = cases.normal((2, 300, 3), 31)
    stats = {"u": {"mean": torch.tensor([0.1, 0.2, 0.3]), "std": torch.tensor([1.1, 0.9, 1.2])},
             "der": {"mean": torch.tensor([0.0, 0.1, -0.1]), "std": torch.tensor([0.5, 0.6, 0.7])},
             "start_time": {"mean": 0.5, "std": 0.3}, "time_diffs": {"mean": 0.1, "std": 0.05}}
    tv = torch.linspace(0, 1, 21)
    idx = list(range(11))
    got = model.autoregressive_predict(u0.to(DEV), idx, tv, stats, "time_der", latent_tokens_coord=lat.to(DEV),
                                       fixed_coord=x.to(DEV), use_cuda_graph=use_graph)
    cfg = orc.make_config(c["latent"], c["magno"], c["transformer"], c["attn"])
    p = {k: v.detach().double().cpu() for k, v in model.state_dict().items()}
    st64 = {k: ({kk: (vv.double() if torch.is_tensor(vv) else vv) for kk, vv in v.items()}) for k, v in stats.items()}
    want = orc.autoregressive_predict(p, cfg, u0.double(), idx, tv.double(), st64, "time_der", lat.double(), x.double())
    for i in range(10):
        assert rel_l2(got[:, i], want[:, i]) < 2e-5, i          # errors compound over the 10 autoregressive steps


@pytest.mark.parametrize("mode", ["output", "residual", "time_der"])
def test_rollout_matches_reference_golden_gpu(mode):
    """CUDA rollout vs the reference's own GAOT.autoregressive_predict output (tests/golden/rollout.npz)."""
    c, x0, tv, stats = cases.rollout_case()
    g = load_golden("rollout")
    model = build_model(c).eval()
    model.load_state_dict({k: v.to(DEV) for k, v in golden_params(g).items()})
    lat, x, _, _ = cases.case_inputs(cases.ROLLOUT["base"])
    got = model.autoregressive_predict(x0.to(DEV), cases.ROLLOUT["time_indices"], tv, stats, mode,
                                       latent_tokens_coord=lat.to(DEV), fixed_coord=x.to(DEV))
    want = torch.from_numpy(g["pred." + mode])
    for i in range(want.shape[1]):
        assert rel_l2(got[:, i], want[:, i]) < 1e-5, (mode, i)


def test_input_prefetcher_feeds_static_buffers():
    """Pinned host batch -> staging on the copy stream -> static buffers on the compute stream (trainer.InputPrefetcher)."""
    from gaot_b200.trainer import InputPrefetcher
    a_s, b_s = torch.zeros(4, 100, 3, device=DEV), torch.zeros(4, 100, 1, device=DEV)
    feeder = InputPrefetcher([a_s, b_s])
    for k in range(3):
        a_h, b_h = torch.full((4, 100, 3), float(k + 1)).pin_memory(), torch.full((4, 100, 1), -float(k + 1)).pin_memory()
        feeder.prefetch([a_h, b_h])
        feeder.commit()
        torch.cuda.synchronize()
        assert torch.equal(a_s.cpu(), a_h) and torch.equal(b_s.cpu(), b_h)
