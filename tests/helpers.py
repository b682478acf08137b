"""Shared test helpers: golden loading, error metrics, model construction from oracle/cases.py."""
import os
from types import SimpleNamespace

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    return {k: z[k] for k in z.files}


def golden_shapes(g):
    return {str(k): [int(t) for t in str(s).split(",")] for k, s in zip(g["keys"], g["shapes"])}


def golden_params(g):
    return {str(k): torch.from_numpy(g["param." + str(k)]) for k in g["keys"]}


def golden_graphs(g, tag, vx):
    """fx: list[scale] of dicts; vx: list[sample][scale]."""
    out = {}
    for k in g:
        if k.startswith(f"graph.{tag}.") and k.endswith(".index"):
            _, _, b, s, _ = k.split(".")
            out.setdefault(int(b), {})[int(s)] = {
                "neighbors_index": torch.from_numpy(g[k].astype(np.int64)),
                "neighbors_row_splits": torch.from_numpy(g[k.replace(".index", ".splits")].astype(np.int64))}
    per = [[out[b][s] for s in sorted(out[b])] for b in sorted(out)]
    return per if vx else per[0]


def rel_l2(a, b):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def grad_scale(g, keys):
    """Largest reference gradient norm of a fixture: structurally zero gradients (e.g. the key bias of dot-product
    attention, to which a row softmax is invariant) are pure rounding noise and are compared against this scale."""
    return max(float(torch.as_tensor(g["grad." + str(k)]).double().norm()) for k in keys if ("grad." + str(k)) in g)


def grad_rel(a, b, scale):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    return float((a - b).norm() / max(float(b.norm()), 1e-6 * scale))


def graphs_to(graphs, device):
    if isinstance(graphs, dict):
        return {k: v.to(device) for k, v in graphs.items()}
    return [graphs_to(g, device) for g in graphs]


def build_model(case: dict, device="cuda"):
    """gaot_b200.model.GAOT for a case of oracle/cases.py (the same hyper-parameters the golden generator
    gave the reference)."""
    from gaot_b200.model import GAOT, AttentionConfig, MAGNOConfig, TransformerConfig
    t = dict(case["transformer"])
    t["attn_config"] = AttentionConfig(**case["attn"])
    cfg = SimpleNamespace(latent_tokens_size=tuple(case["latent"]),
                          args=SimpleNamespace(magno=MAGNOConfig(**case["magno"]), transformer=TransformerConfig(**t)))
    return GAOT(case["cin"], case["cout"], cfg).to(device)
