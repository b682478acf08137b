"""Shared test helpers: golden loading and error metrics."""
import os

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
