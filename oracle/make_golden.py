"""Golden-vector generator: runs the UNMODIFIED reference `src/model` package (imported from
/root/reference, CPU) on the deterministic cases of `oracle/cases.py` and writes small fixtures to
`tests/golden/`.  Run in the build container only (the GPU box has no /root/reference):

    python oracle/make_golden.py [case ...]

The reference needs three wheels that are absent here (torch_scatter, omegaconf,
rotary_embedding_torch); `oracle/ref_shims/` provides import stand-ins with the documented
semantics (see its README).  Truth values are computed by the reference in fp64
(`torch.set_default_dtype(float64)` because gemb.py:146 ignores the input dtype) and in fp32.
"""
from __future__ import annotations

import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "ref_shims"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

from src.model.gaot import GAOT  # noqa: E402  (the reference)
from src.model.layers.magno import MAGNOConfig  # noqa: E402
from src.model.layers.attn import TransformerConfig, AttentionConfig  # noqa: E402
from src.model.layers.utils.neighbor_search import NeighborSearch  # noqa: E402

from oracle import cases, gaot_oracle as orc  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def ref_config(c):
    t = dict(c["transformer"])
    t["attn_config"] = AttentionConfig(**c["attn"])
    return SimpleNamespace(latent_tokens_size=tuple(c["latent"]),
                           args=SimpleNamespace(magno=MAGNOConfig(**c["magno"]), transformer=TransformerConfig(**t)))


def run_reference(name, dtype):
    c = cases.CASES[name]
    torch.set_default_dtype(dtype)
    try:
        torch.manual_seed(42)
        model = GAOT(c["cin"], c["cout"], ref_config(c))
        shapes = {k: list(v.shape) for k, v in model.state_dict().items()}
        params = cases.synth_params(shapes, 7 + c["seed"])
        model.load_state_dict({k: v.to(dtype) for k, v in params.items()})
        model.train()
        lat, x, pn, tgt = [t.to(dtype) for t in cases.case_inputs(name)]
        enc_n = dec_n = None
        graphs = {}
        if c.get("vx"):
            ns = NeighborSearch("native")
            r = c["magno"]["radius"]
            enc_n = [[ns(x[b].float(), lat.float(), r * s) for s in model.encoder.scales] for b in range(c["B"])]
            dec_n = [[ns(lat.float(), x[b].float(), r * s) for s in model.encoder.scales] for b in range(c["B"])]
        inter = {}
        hooks = [model.encoder.register_forward_hook(lambda m, i, o: inter.__setitem__("enc_out", o.detach())),
                 model.processor.register_forward_hook(lambda m, i, o: inter.__setitem__("proc_out", o.detach()))]
        cond = cases.case_condition(name)
        cond = cond.to(dtype) if cond is not None else None
        out = model(lat, x, pn, encoder_nbrs=enc_n, decoder_nbrs=dec_n, condition=cond)
        loss = torch.nn.functional.mse_loss(out, tgt)
        loss.backward()
        for h in hooks:
            h.remove()
        grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
        if not c.get("vx"):
            graphs["enc"] = model.encoder.neighbor_cache[next(iter(model.encoder.neighbor_cache))]
            graphs["dec"] = model.decoder.neighbor_cache[next(iter(model.decoder.neighbor_cache))]
        else:
            graphs["enc"], graphs["dec"] = enc_n, dec_n
        return dict(shapes=shapes, params=params, out=out.detach(), loss=loss.detach(), grads=grads,
                    graphs=graphs, inter=inter)
    finally:
        torch.set_default_dtype(torch.float32)


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def grel(a, b, scale):
    """gradient error; structurally zero gradients (pure rounding noise) are measured against the model's gradient scale"""
    return float((a.double() - b.double()).norm() / max(float(b.double().norm()), 1e-6 * scale))


def check_oracle(name, ref64):
    """The oracle restatement must reproduce the reference in fp64 to ~1e-12 (same arithmetic)."""
    c = cases.CASES[name]
    cfg = orc.make_config(c["latent"], c["magno"], c["transformer"], c["attn"])
    lat, x, pn, tgt = [t.double() for t in cases.case_inputs(name)]
    p = {k: v.double().requires_grad_(True) for k, v in ref64["params"].items()}
    enc_n = dec_n = None
    if c.get("vx"):
        enc_n, dec_n = ref64["graphs"]["enc"], ref64["graphs"]["dec"]
    cond = cases.case_condition(name)
    cond = cond.double() if cond is not None else None
    loss = orc.train_step_loss(p, cfg, lat, x, pn, tgt, enc_n, dec_n, cond)
    loss.backward()
    e_loss = abs(float(loss.detach()) - float(ref64["loss"])) / abs(float(ref64["loss"]))
    gscale = max(float(v.double().norm()) for v in ref64["grads"].values())
    e_g = max(grel(p[k].grad, ref64["grads"][k], gscale) for k in p)
    out = orc.gaot_forward({k: v.detach() for k, v in p.items()}, cfg, lat, x, pn, None, enc_n, dec_n, cond)
    e_out = rel(out, ref64["out"])
    print(f"  oracle-vs-reference fp64: out {e_out:.2e} loss {e_loss:.2e} worst grad {e_g:.2e}")
    assert e_out < 1e-10 and e_loss < 1e-10 and e_g < 1e-8, "oracle restatement diverges from the reference"
    # fp32 noise floor of the same arithmetic on the same graphs (sets per-tensor test tolerances)
    p32 = {k: v.detach().clone().float().requires_grad_(True) for k, v in ref64["params"].items()}
    if enc_n is None:
        enc_n, dec_n = ref64["graphs"]["enc"], ref64["graphs"]["dec"]
    l32 = orc.train_step_loss(p32, cfg, lat.float(), x.float(), pn.float(), tgt.float(), enc_n, dec_n,
                              cond.float() if cond is not None else None)
    l32.backward()
    ref64["gerr32"] = {k: grel(p32[k].grad, ref64["grads"][k], gscale) for k in p32}
    print(f"  fp32 noise floor (oracle fp32 vs fp64, same graph): worst grad {max(ref64['gerr32'].values()):.2e}")
    # graphs: canonical definition == reference `native` on these inputs
    if not c.get("vx"):
        mc = cfg.args.magno
        for tag, (dd, qq) in {"enc": (x.float(), lat.float()), "dec": (lat.float(), x.float())}.items():
            for s, g in zip(mc.scales, ref64["graphs"][tag]):
                o = orc.radius_csr(dd, qq, mc.radius * s)
                assert torch.equal(o["neighbors_index"], g["neighbors_index"]), (name, tag)
                assert torch.equal(o["neighbors_row_splits"], g["neighbors_row_splits"]), (name, tag)


def graph_arrays(graphs, vx):
    out = {}
    for tag in ("enc", "dec"):
        gl = graphs[tag]
        if not vx:
            gl = [gl]
        for b, per_scale in enumerate(gl):
            for s, g in enumerate(per_scale):
                out[f"graph.{tag}.{b}.{s}.index"] = g["neighbors_index"].numpy().astype(np.int32)
                out[f"graph.{tag}.{b}.{s}.splits"] = g["neighbors_row_splits"].numpy().astype(np.int32)
    return out


def rollout_golden():
    """Reference GAOT.autoregressive_predict (gaot.py:307-476) in fp64 on the rollout fixture, every stepper mode;
    the oracle's restatement must match it to 1e-10."""
    c, x0, tv, stats = cases.rollout_case()
    torch.set_default_dtype(torch.float64)
    try:
        torch.manual_seed(42)
        model = GAOT(c["cin"], c["cout"], ref_config(c)).eval()
        shapes = {k: list(v.shape) for k, v in model.state_dict().items()}
        params = cases.synth_params(shapes, 7 + cases.ROLLOUT["seed"])
        model.load_state_dict({k: v.double() for k, v in params.items()})
        lat, x, _, _ = cases.case_inputs(cases.ROLLOUT["base"])
        st64 = {k: {kk: (vv.double() if torch.is_tensor(vv) else vv) for kk, vv in v.items()} for k, v in stats.items()}
        cfg = orc.make_config(c["latent"], c["magno"], c["transformer"], c["attn"])
        arrs = {"keys": np.array(list(shapes.keys())), "shapes": np.array([",".join(map(str, s)) for s in shapes.values()])}
        for k, v in params.items():
            arrs["param." + k] = v.numpy().astype(np.float32)
        ti = np.array(cases.ROLLOUT["time_indices"])
        for mode in ("output", "residual", "time_der"):
            # a dataset sample also carries its own time channels after [u | c]; the reference ignores them
            xb = torch.cat([x0.double(), torch.zeros(x0.shape[0], x0.shape[1], 2, dtype=torch.float64)], -1)
            ref = model.autoregressive_predict(xb, ti, tv.double().numpy(), st64, stepper_mode=mode,
                                               latent_tokens_coord=lat.double(), fixed_coord=x.double())
            mine = orc.autoregressive_predict({k: v.double() for k, v in params.items()}, cfg, xb, ti, tv.double(), st64, mode,
                                              lat.double(), x.double())
            e = rel(mine, ref)
            print(f"[rollout {mode}] oracle-vs-reference fp64: {e:.2e}, |pred| {float(ref.norm()):.3f}")
            assert e < 1e-10
            arrs["pred." + mode] = ref.numpy().astype(np.float32)
        np.savez_compressed(os.path.join(OUT, "rollout.npz"), **arrs)
    finally:
        torch.set_default_dtype(torch.float32)


def _load_ref_file(name, path, package=None):
    """Import ONE reference source file without running its package __init__ (src/datasets/__init__.py needs xarray,
    src/utils/__init__.py needs matplotlib - both absent here)."""
    import importlib.util
    import types
    if package and package not in sys.modules:
        pkg = types.ModuleType(package)
        pkg.__path__ = []
        sys.modules[package] = pkg
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def datasets_golden():
    """Reference GraphBuilder.build_graphs_for_split (graph_builder.py:28-87, with utils/scaling.rescale) and
    DynamicPairDataset (data_utils.py:73-235) run on the dataset fixtures of oracle/cases.py."""
    _load_ref_file("src.utils.scaling", "/root/reference/src/utils/scaling.py", "src.utils")
    gb = _load_ref_file("src.datasets.graph_builder", "/root/reference/src/datasets/graph_builder.py", "src.datasets")
    du = _load_ref_file("src.datasets.data_utils", "/root/reference/src/datasets/data_utils.py", "src.datasets")
    arrs = {}
    x_data, lat, r, scales = cases.graph_builder_case()
    enc, dec = gb.GraphBuilder("native").build_graphs_for_split(x_data, lat, r, scales)
    for tag, graphs in (("enc", enc), ("dec", dec)):
        for b, per_scale in enumerate(graphs):
            for s_, g in enumerate(per_scale):
                arrs[f"gb.{tag}.{b}.{s_}.index"] = g["neighbors_index"].numpy().astype(np.int32)
                arrs[f"gb.{tag}.{b}.{s_}.splits"] = g["neighbors_row_splits"].numpy().astype(np.int32)
    print(f"[graph builder] {len(enc)} samples x {len(scales)} scales, E(enc) = "
          f"{[int(g[0]['neighbors_index'].numel()) for g in enc]}")
    u, c, tv, x, stats = cases.pair_dataset_case()
    for mode in ("output", "residual", "time_der"):
        for vx in (False, True):
            ds = du.DynamicPairDataset(u, c, tv, None, max_time_diff=6, time_step=2, stepper_mode=mode, stats=stats,
                                       x_data=x if vx else None, is_variable_coords=vx)
            tag = f"pairs.{mode}.{'vx' if vx else 'fx'}"
            arrs[tag + ".len"] = np.int64(len(ds))
            arrs[tag + ".t_in"], arrs[tag + ".t_out"] = ds.t_in_indices, ds.t_out_indices
            pick = list(range(0, len(ds), 5))
            items = [ds[i] for i in pick]
            arrs[tag + ".pick"] = np.array(pick)
            arrs[tag + ".input"] = torch.stack([it[0] for it in items]).numpy()
            arrs[tag + ".target"] = torch.stack([it[1] for it in items]).numpy()
            if vx:
                arrs[tag + ".x"] = torch.stack([it[2] for it in items]).numpy()
    ds = du.DynamicPairDataset(u, None, tv, None, max_time_diff=14, time_step=1, stepper_mode="residual", stats=None,
                               use_time_norm=False)
    arrs["pairs.nostats.len"] = np.int64(len(ds))
    arrs["pairs.nostats.input"] = torch.stack([ds[i][0] for i in range(0, len(ds), 7)]).numpy()
    arrs["pairs.nostats.target"] = torch.stack([ds[i][1] for i in range(0, len(ds), 7)]).numpy()
    print(f"[pair dataset] {int(arrs['pairs.output.fx.len'])} pairs per configuration")
    np.savez_compressed(os.path.join(OUT, "datasets.npz"), **arrs)


def agno_golden():
    """AGNO options outside the MAGNO call sites, against the reference's own AGNO module (agno.py:67-273): the
    kernel-only transform (f_y=None) and the f_y transform, with cosine attention (sum) and without (mean).  The
    integration-weights branch (agno.py:252-262) cannot run in the reference for C > 1 (an [E] vector is multiplied into
    [E, C] without unsqueezing), so `weights=` is pinned by the oracle restatement of its documented meaning only."""
    from src.model.layers.agno import AGNO
    y, x, f_y, wts, layers, radius = cases.agno_case()
    nbrs = NeighborSearch("native")(y, x, radius)
    arrs = {"index": nbrs["neighbors_index"].numpy(), "splits": nbrs["neighbors_row_splits"].numpy()}
    torch.set_default_dtype(torch.float64)
    try:
        for tag, use_attn in (("attn", True), ("mean", False)):
            m = AGNO(channel_mlp_layers=layers, use_attn=use_attn, attention_type="cosine", coord_dim=2, use_torch_scatter=True)
            shapes = {k: list(v.shape) for k, v in m.state_dict().items()}
            params = cases.synth_params(shapes, 710 + int(use_attn))
            m.load_state_dict({k: v.double() for k, v in params.items()})
            y64, x64, f64 = y.double(), x.double(), f_y.double()
            konly = m(y64, nbrs, x64, f_y=None)
            full = m(y64, nbrs, x64, f_y=f64)
            p = {"agno." + k: v.double() for k, v in params.items()}
            o1 = orc.agno_kernel_only(p, "agno", y64, nbrs, x64, use_attn=use_attn, coord_dim=2)
            o2 = orc.agno(p, "agno", y64, nbrs, x64, f64, use_attn=use_attn, coord_dim=2)
            assert rel(o1, konly) < 1e-10 and rel(o2, full) < 1e-10, (tag, rel(o1, konly), rel(o2, full))
            for k, v in params.items():
                arrs[f"{tag}.param.{k}"] = v.numpy()
            arrs[f"{tag}.kernel_only"] = konly.detach().numpy().astype(np.float32)
            arrs[f"{tag}.full"] = full.detach().numpy().astype(np.float32)
            # integration weights: oracle restatement only (see above)
            arrs[f"{tag}.kernel_only_weighted"] = orc.agno_kernel_only(p, "agno", y64, nbrs, x64, use_attn=use_attn, coord_dim=2,
                                                                      weights=wts.double()).numpy().astype(np.float32)
            arrs[f"{tag}.full_weighted"] = orc.agno_weighted(p, "agno", y64, nbrs, x64, f64, wts.double(), use_attn=use_attn,
                                                             coord_dim=2).numpy().astype(np.float32)
    finally:
        torch.set_default_dtype(torch.float32)
    print(f"[agno options] E = {arrs['index'].size}, kernel-only / f_y transforms match the reference AGNO")
    np.savez_compressed(os.path.join(OUT, "agno.npz"), **arrs)


def main():
    os.makedirs(OUT, exist_ok=True)
    only = sys.argv[1:]                                   # optional: regenerate just these model cases
    if not only or "agno" in only:
        agno_golden()
    if not only or "rollout" in only:
        rollout_golden()
    if not only or "datasets" in only:
        datasets_golden()
    if only and all(o in ("rollout", "datasets", "agno") for o in only):
        return
    for name, c in cases.CASES.items():
        if only and name not in only:
            continue
        print(f"[case {name}]")
        r64 = run_reference(name, torch.float64)
        r32 = run_reference(name, torch.float32)
        check_oracle(name, r64)
        print(f"  reference fp32 vs fp64: out {rel(r32['out'], r64['out']):.2e}, worst grad "
              f"{max(rel(r32['grads'][k], r64['grads'][k]) for k in r64['grads']):.2e}")
        arrs = {"loss64": np.float64(r64["loss"]), "loss32": np.float32(r32["loss"])}
        arrs["gerr32"] = np.array([r64["gerr32"][k] for k in r64["shapes"]])
        arrs["keys"] = np.array(list(r64["shapes"].keys()))
        arrs["shapes"] = np.array([",".join(map(str, s)) for s in r64["shapes"].values()])
        if c.get("summary_only"):
            # params are regenerated from (shapes, seed) by cases.synth_params; keep only summaries
            arrs["out_sub"] = r64["out"][:, ::37, :].numpy().astype(np.float32)
            arrs["out_norm"] = np.float64(r64["out"].norm())
            arrs["enc_sub"] = r64["inter"]["enc_out"][:, ::41, ::5].numpy().astype(np.float32)
            arrs["proc_sub"] = r64["inter"]["proc_out"][:, ::13, ::7].numpy().astype(np.float32)
            arrs["grad_norms"] = np.array([float(r64["grads"][k].norm()) for k in r64["shapes"]])
            for k in ("encoder.geoembed.mlp.0.weight", "decoder.projection.fcs.0.weight",
                      "encoder.agno.channel_mlp.fcs.0.weight", "decoder.agno.channel_mlp.fcs.3.bias",
                      "processor.middle_layer.attn_norm.weight", "encoder.lifting.fcs.0.weight"):
                arrs["grad." + k] = r64["grads"][k].numpy().astype(np.float32)
            g = r64["graphs"]
            arrs["enc_E"] = np.int64(g["enc"][0]["neighbors_index"].numel())
            arrs["enc_index_sum"] = np.int64(g["enc"][0]["neighbors_index"].sum())
            arrs["enc_splits_sub"] = g["enc"][0]["neighbors_row_splits"][::64].numpy()
            arrs["dec_index_sum"] = np.int64(g["dec"][0]["neighbors_index"].sum())
            arrs["dec_splits_sub"] = g["dec"][0]["neighbors_row_splits"][::128].numpy()
        else:
            for k, v in r64["params"].items():
                arrs["param." + k] = v.numpy().astype(np.float32)
            for k, v in r64["grads"].items():
                arrs["grad." + k] = v.numpy().astype(np.float32)
            arrs["out"] = r64["out"].numpy().astype(np.float32)
            arrs["out32"] = r32["out"].numpy().astype(np.float32)
            arrs["enc_out"] = r64["inter"]["enc_out"].numpy().astype(np.float32)
            arrs["proc_out"] = r64["inter"]["proc_out"].numpy().astype(np.float32)
            arrs.update(graph_arrays(r64["graphs"], bool(c.get("vx"))))
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **arrs)

    if only:
        return
    # neighbour search: every reference back-end that runs here, on every nbr case
    arrs = {}
    for nm in cases.NBR_CASES:
        data, q, r = cases.nbr_case(nm)
        nat = NeighborSearch("native")(data, q, r)
        chk = NeighborSearch("chunked")(data, q, r)
        assert torch.equal(nat["neighbors_index"], chk["neighbors_index"].to(nat["neighbors_index"].dtype))
        o = orc.radius_csr(data, q, r)
        assert torch.equal(o["neighbors_index"], nat["neighbors_index"]), nm
        assert torch.equal(o["neighbors_row_splits"], nat["neighbors_row_splits"]), nm
        if data.shape[1] == 2 and nm != "empty":
            grd = NeighborSearch("grid")(data, q, r)   # pure-python 2-D grid: same sets, cell order
            rs = grd["neighbors_row_splits"]
            assert torch.equal(rs.long(), nat["neighbors_row_splits"]), nm
            for m in range(q.shape[0]):
                a = torch.sort(grd["neighbors_index"][rs[m]:rs[m + 1]].long()).values
                assert torch.equal(a, nat["neighbors_index"][rs[m]:rs[m + 1]]), (nm, m)
        arrs[f"{nm}.index"] = nat["neighbors_index"].numpy().astype(np.int32)
        arrs[f"{nm}.splits"] = nat["neighbors_row_splits"].numpy().astype(np.int32)
        print(f"[nbr {nm}] E={nat['neighbors_index'].numel()} ok (native == chunked == oracle canon)")
    np.savez_compressed(os.path.join(OUT, "neighbors.npz"), **arrs)
    print("golden vectors written to", OUT)


if __name__ == "__main__":
    main()
