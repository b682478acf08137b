"""Deterministic synthetic inputs shared by the golden generator, the tests, smoke() and bench.py.

TEST / BENCH INFRASTRUCTURE.  Nothing here touches /root/reference; everything is a pure function of
a seed so the GPU box can regenerate the same tensors the golden vectors were produced from
(coords in [-1,1]^d, latent grid = meshgrid(linspace(-1,1,H),..., indexing='ij')
as reference src/datasets/data_processor.py:289-307, features ~ N(0,1))."""
from __future__ import annotations

import math
from typing import Dict, List

import torch


from gaot_b200.synth import clustered_points, latent_grid, normal, uniform_points  # noqa: E402,F401


def synth_params(shapes: Dict[str, List[int]], seed: int) -> Dict[str, torch.Tensor]:
    """Deterministic parameter values for a given {state_dict key: shape} table (key order matters):
    matrices U(-1,1)/sqrt(fan_in), biases U(-1,1)*0.1, norm gains 1 + 0.1*U(-1,1)."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k, shp in shapes.items():
        u = torch.rand(*shp, generator=g, dtype=torch.float32) * 2.0 - 1.0
        if k.endswith("norm.weight"):
            out[k] = 1.0 + 0.1 * u
        elif k.endswith(".bias"):
            out[k] = 0.1 * u
        else:
            fan_in = int(math.prod(shp[1:])) if len(shp) > 1 else shp[0]
            out[k] = u / math.sqrt(max(fan_in, 1))
    return out


# ---- model cases (hyper-parameters only; tensors come from the generators above) ----------------
CASES = {
    # tiny 2-D fx model: head_dim 32, Dp == Dh (Identity in/out proj)
    "tiny2d": dict(latent=(8, 8), n=300, d=2, B=2, cin=2, cout=3, seed=1,
                   magno=dict(coord_dim=2, radius=0.35, hidden_size=16, mlp_layers=3, lifting_channels=16,
                              neighbor_search_method="native"),
                   transformer=dict(patch_size=2, hidden_size=64, num_layers=3, ffn_multiplier=2),
                   attn=dict(num_heads=2, num_kv_heads=2)),
    # tiny 3-D fx, two scales with learned scale weights, Dp=48 != Dh=64 (input/output proj), 5 layers
    "tiny3d": dict(latent=(4, 4, 4), n=400, d=3, B=2, cin=3, cout=1, seed=2,
                   magno=dict(coord_dim=3, radius=0.6, hidden_size=16, mlp_layers=2, lifting_channels=6,
                              scales=[1.0, 1.5], use_scale_weights=True, neighbor_search_method="native"),
                   transformer=dict(patch_size=2, hidden_size=64, num_layers=5, ffn_multiplier=2),
                   attn=dict(num_heads=2, num_kv_heads=2)),
    # vx: per-sample meshes and precomputed per-sample graphs
    "vx2d": dict(latent=(8, 8), n=250, d=2, B=3, cin=3, cout=1, seed=3, vx=True,
                 magno=dict(coord_dim=2, radius=0.4, hidden_size=16, mlp_layers=3, lifting_channels=16,
                            precompute_edges=True, neighbor_search_method="native"),
                 transformer=dict(patch_size=2, hidden_size=64, num_layers=2, ffn_multiplier=2),
                 attn=dict(num_heads=2, num_kv_heads=2)),
    # no attention (mean reduce), no geo-embedding
    "plain2d": dict(latent=(8, 8), n=300, d=2, B=2, cin=1, cout=1, seed=4,
                    magno=dict(coord_dim=2, radius=0.35, hidden_size=16, mlp_layers=3, lifting_channels=16,
                               use_attention=False, use_geoembed=False, neighbor_search_method="native"),
                    transformer=dict(patch_size=2, hidden_size=64, num_layers=1, ffn_multiplier=2),
                    attn=dict(num_heads=2, num_kv_heads=2)),
    # learned dot-product attention weights (agno.py:105-110, 214-217) instead of the parameter-free cosine attention
    "dot2d": dict(latent=(8, 8), n=300, d=2, B=2, cin=2, cout=1, seed=5,
                  magno=dict(coord_dim=2, radius=0.35, hidden_size=16, mlp_layers=2, lifting_channels=16,
                             attention_type="dot_product", neighbor_search_method="native"),
                  transformer=dict(patch_size=2, hidden_size=64, num_layers=1, ffn_multiplier=2),
                  attn=dict(num_heads=2, num_kv_heads=2)),
    # sinusoidal node position encoding of the kernel / attention coordinates (gemb.py:12-34, magno.py:161-162, 323-325)
    "npe2d": dict(latent=(8, 8), n=300, d=2, B=2, cin=1, cout=2, seed=6,
                  magno=dict(coord_dim=2, radius=0.35, hidden_size=16, mlp_layers=2, lifting_channels=16,
                             node_embedding=True, neighbor_search_method="native"),
                  transformer=dict(patch_size=2, hidden_size=64, num_layers=1, ffn_multiplier=2),
                  attn=dict(num_heads=2, num_kv_heads=2)),
    # pointnet geometry embedding with max pooling (gemb.py:60-70, 173-228)
    "pnet2d": dict(latent=(8, 8), n=300, d=2, B=2, cin=1, cout=1, seed=7,
                   magno=dict(coord_dim=2, radius=0.3, hidden_size=16, mlp_layers=2, lifting_channels=16,
                              embedding_method="pointnet", pooling="max", neighbor_search_method="native"),
                   transformer=dict(patch_size=2, hidden_size=64, num_layers=1, ffn_multiplier=2),
                   attn=dict(num_heads=2, num_kv_heads=2)),
    # time-conditioned affine corrections in attention and FFN (attn.py:70-71, 89-90, 145-146, 153-154; mlp.py:74-124)
    "cond2d": dict(latent=(8, 8), n=300, d=2, B=3, cin=2, cout=1, seed=8, cond=True,
                   magno=dict(coord_dim=2, radius=0.35, hidden_size=16, mlp_layers=2, lifting_channels=16,
                              neighbor_search_method="native"),
                   transformer=dict(patch_size=2, hidden_size=64, num_layers=3, ffn_multiplier=2),
                   attn=dict(num_heads=2, num_kv_heads=2, use_conditional_norm=True)),
    # the optional paths together, in vx mode: per-sample meshes/graphs + dot-product attention + pointnet (mean pooling)
    # + node position encoding
    "vxmix": dict(latent=(8, 8), n=220, d=2, B=3, cin=2, cout=2, seed=9, vx=True,
                  magno=dict(coord_dim=2, radius=0.4, hidden_size=16, mlp_layers=2, lifting_channels=16,
                             attention_type="dot_product", embedding_method="pointnet", pooling="mean", node_embedding=True,
                             precompute_edges=True, neighbor_search_method="native"),
                  transformer=dict(patch_size=2, hidden_size=64, num_layers=2, ffn_multiplier=2),
                  attn=dict(num_heads=2, num_kv_heads=2)),
    # transform_type 'linear_kernelonly' (agno.py:229-247: same arithmetic as 'linear' once f_y is given), mean reduce
    "konly2d": dict(latent=(8, 8), n=300, d=2, B=2, cin=2, cout=2, seed=10,
                    magno=dict(coord_dim=2, radius=0.35, hidden_size=16, mlp_layers=2, lifting_channels=16,
                               transform_type="linear_kernelonly", use_attention=False, neighbor_search_method="native"),
                    transformer=dict(patch_size=2, hidden_size=64, num_layers=1, ffn_multiplier=2),
                    attn=dict(num_heads=2, num_kv_heads=2)),
    # rotary position embedding instead of the absolute table (attn.py:75-76, 106-108, gaot.py:217-218); GQA 4 / 2 heads
    "rope2d": dict(latent=(8, 8), n=300, d=2, B=2, cin=2, cout=1, seed=11,
                   magno=dict(coord_dim=2, radius=0.35, hidden_size=16, mlp_layers=2, lifting_channels=16,
                              neighbor_search_method="native"),
                   transformer=dict(patch_size=2, hidden_size=64, num_layers=2, ffn_multiplier=2, positional_embedding="rope"),
                   attn=dict(num_heads=4, num_kv_heads=2)),
    # BASELINE cfg-2 geometry (Poisson-Gauss hyper-parameters, config/examples/time_indep/poisson_gauss.json)
    "cfg2": dict(latent=(64, 64), n=9216, d=2, B=2, cin=1, cout=1, seed=0, summary_only=True,
                 magno=dict(coord_dim=2, radius=0.033, hidden_size=64, mlp_layers=3, lifting_channels=64,
                            neighbor_search_method="native"),
                 transformer=dict(patch_size=2, hidden_size=256, num_layers=3, ffn_multiplier=4),
                 attn=dict(num_heads=8, num_kv_heads=8)),
}


# autoregressive rollout fixture (gaot.py:307-476): tiny2d geometry, u(3) + c(1) + t0 + dt input channels, all stepper modes
ROLLOUT = dict(base="tiny2d", u_dim=3, c_dim=1, B=2, seed=31, time_indices=[0, 2, 3, 6, 7], n_times=11)


def rollout_case():
    """(case dict, x_batch [B,N,u+c], t_values, stats) for the rollout fixture; stats in the layout the reference's
    sequential trainer passes (per-channel tensors for u / res / der / c, floats for the time statistics)."""
    r = ROLLOUT
    c = dict(CASES[r["base"]], cin=r["u_dim"] + r["c_dim"] + 2, cout=r["u_dim"], B=r["B"])
    x0 = normal((r["B"], c["n"], r["u_dim"] + r["c_dim"]), r["seed"])
    tv = torch.linspace(0, 1, r["n_times"])
    stats = {"u": {"mean": torch.tensor([0.1, 0.2, 0.3]), "std": torch.tensor([1.1, 0.9, 1.2])},
             "c": {"mean": torch.tensor([0.4]), "std": torch.tensor([2.0])},
             "der": {"mean": torch.tensor([0.0, 0.1, -0.1]), "std": torch.tensor([0.5, 0.6, 0.7])},
             "res": {"mean": torch.tensor([0.05, -0.1, 0.2]), "std": torch.tensor([0.3, 0.4, 0.2])},
             "start_time": {"mean": 0.5, "std": 0.3}, "time_diffs": {"mean": 0.1, "std": 0.05}}
    return c, x0, tv, stats


def case_inputs(name: str):
    c = CASES[name]
    lat = latent_grid(c["latent"])
    if c.get("vx"):
        x = torch.stack([clustered_points(c["n"], c["d"], 100 + c["seed"] * 10 + b) for b in range(c["B"])])
    else:
        x = uniform_points(c["n"], c["d"], c["seed"])
    pn = normal((c["B"], c["n"], c["cin"]), 1000 + c["seed"])
    tgt = normal((c["B"], c["n"], c["cout"]), 2000 + c["seed"])
    return lat, x, pn, tgt


def case_condition(name: str):
    """[B, 1] conditioning values (time differences in (0, 1)) for cases with `cond`, else None."""
    c = CASES[name]
    if not c.get("cond"):
        return None
    return torch.rand(c["B"], 1, generator=torch.Generator().manual_seed(3000 + c["seed"]), dtype=torch.float32)


# ---- neighbour-search cases ---------------------------------------------------------------------
def nbr_case(name: str):
    """(data, queries, radius).  'lattice*' cases sit on dyadic coordinates so that distances equal to
    the radius are exact in fp32 (ties well defined in every back-end); they also contain duplicate
    points and queries with no neighbour."""
    if name == "rand2d":
        return uniform_points(2000, 2, 11), latent_grid((16, 16)), 0.11
    if name == "rand3d":
        return uniform_points(3000, 3, 12), latent_grid((6, 6, 6)), 0.3
    if name == "dec2d":
        return latent_grid((16, 16)), uniform_points(1500, 2, 13), 0.15
    if name == "lattice2d":
        ax = torch.arange(-8, 9, dtype=torch.float32) * 0.125
        pts = torch.stack(torch.meshgrid(ax, ax, indexing="ij"), -1).reshape(-1, 2)
        pts = torch.cat([pts, pts[::7]], dim=0)                     # duplicates
        q = torch.cat([pts[::3] * 1.0, torch.tensor([[5.0, 5.0], [-3.0, 0.0]])], dim=0)  # + empty rows
        return pts.contiguous(), q.contiguous(), 0.25
    if name == "lattice3d":
        ax = torch.arange(-3, 4, dtype=torch.float32) * 0.25
        pts = torch.stack(torch.meshgrid(ax, ax, ax, indexing="ij"), -1).reshape(-1, 3)
        q = torch.cat([pts[::5], torch.tensor([[4.0, 4.0, 4.0]])], dim=0)
        return pts.contiguous(), q.contiguous(), 0.5
    if name == "empty":
        return uniform_points(50, 2, 14), uniform_points(20, 2, 15) + 10.0, 0.1
    raise KeyError(name)


NBR_CASES = ["rand2d", "rand3d", "dec2d", "lattice2d", "lattice3d", "empty"]


# ---- dataset-side fixtures (graph builder, time-pair dataset) -----------------------------------------------------
def agno_case():
    """(y [n,2], x [m,2], f_y [2,n,8], weights [n], layers) for the AGNO option fixtures (kernel-only transform, integration
    weights): clustered source points against a 6 x 6 query grid, radius 0.45; one query is moved away so its row is empty."""
    y = clustered_points(120, 2, 700)
    x = latent_grid((6, 6)).clone()
    x[5] = torch.tensor([7.0, 7.0])
    f_y = normal((2, 120, 8), 701)
    wts = torch.rand(120, generator=torch.Generator().manual_seed(702), dtype=torch.float32) * 0.9 + 0.1
    return y, x, f_y, wts, [4, 16, 8], 0.45


def graph_builder_case():
    """(x_data [S, 1, n, 2] UN-scaled sample meshes, latent [M, 2], radius, scales): the builder rescales every sample to
    [-1, 1] by its own min / max (graph_builder.py:60) before searching."""
    S, n = 4, 180
    xs = []
    for b in range(S):
        x = clustered_points(n, 2, 500 + b)
        xs.append(x * torch.tensor([1.0 + 0.5 * b, 2.0]) + torch.tensor([3.0 * b, -1.0]))      # different boxes per sample
    return torch.stack(xs).unsqueeze(1), latent_grid((8, 8)), 0.3, [1.0, 1.5]


def pair_dataset_case():
    """u [S, T, N, U], c [S, T, N, 1], t_values [T], x [S, T, N, 2], stats - small sequential dataset."""
    S, T, N, U = 3, 9, 40, 2
    u = normal((S, T, N, U), 601)
    c = normal((S, T, N, 1), 602)
    x = uniform_points(S * T * N, 2, 603).view(S, T, N, 2)
    tv = torch.linspace(0.0, 1.6, T)
    stats = {"u": {"mean": torch.tensor([0.1, -0.2]), "std": torch.tensor([1.2, 0.8])},
             "c": {"mean": torch.tensor([0.3]), "std": torch.tensor([1.5])},
             "res": {"mean": torch.tensor([0.01, 0.02]), "std": torch.tensor([0.5, 0.6])},
             "der": {"mean": torch.tensor([0.0, 0.1]), "std": torch.tensor([2.0, 3.0])},
             "start_time": {"mean": 0.5, "std": 0.4}, "time_diffs": {"mean": 0.6, "std": 0.3}}
    return u, c, tv, x, stats
