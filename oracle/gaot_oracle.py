"""CPU oracle: a functional restatement of the reference GAOT forward path (torch-CPU + numpy).

TEST INFRASTRUCTURE ONLY - never imported by `gaot_b200/`.  Every function cites the reference
file:line it restates (paths relative to the reference checkout's `src/model/`).  The oracle is
pinned against golden vectors produced by running the UNMODIFIED reference (`oracle/make_golden.py`,
fixtures in `tests/golden/`); `tests/test_oracle_golden.py` is that pin.

Parameters are passed as a flat dict keyed exactly like the reference `state_dict()`
(`encoder.agno.channel_mlp.fcs.0.weight`, ...), so reference checkpoints drive the oracle directly.
Backward comes from torch autograd over these functions (the reference's backward is autograd over
the same arithmetic).  dtype follows the inputs (fp32 or fp64).
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------------
# A1. radius graph  (layers/utils/neighbor_search.py:108-144 `native`, :147-180 `torch_cluster`)
# --------------------------------------------------------------------------------------------

def radius_csr(data, queries, radius: float, inclusive: bool = True,
               max_num_neighbors: Optional[int] = None):
    """Canonical radius graph: row m lists, in ascending j, every j with
    sum_d (data[j,d]-queries[m,d])^2  (<= | <)  r^2, all arithmetic in fp32 without fma contraction
    (products rounded, then added in dimension order), r^2 = fp32(r)*fp32(r).
    `inclusive=True, max_num_neighbors=None` is the reference `native`/`chunked` back-end
    (neighbor_search.py:133-136: cdist <= r, nonzero order = ascending j); `inclusive=False,
    max_num_neighbors=32` is torch_cluster.radius as called at neighbor_search.py:163-165
    (first-k in ascending j on its CUDA path).  Returns the reference dict (int64 tensors)."""
    x = np.asarray(data.detach().cpu().numpy() if torch.is_tensor(data) else data, dtype=np.float32)
    q = np.asarray(queries.detach().cpu().numpy() if torch.is_tensor(queries) else queries, dtype=np.float32)
    r = np.float32(radius)
    r2 = np.float32(r * r)
    m, d = q.shape
    idx_chunks: List[np.ndarray] = []
    counts = np.zeros(m, dtype=np.int64)
    step = max(1, (1 << 24) // max(1, x.shape[0]))
    for s in range(0, m, step):
        qq = q[s:s + step]
        d2 = np.zeros((qq.shape[0], x.shape[0]), dtype=np.float32)
        for k in range(d):
            diff = (x[None, :, k] - qq[:, None, k]).astype(np.float32)
            d2 = (d2 + (diff * diff).astype(np.float32)).astype(np.float32)
        hit = (d2 <= r2) if inclusive else (d2 < r2)
        if max_num_neighbors is not None:
            hit &= (np.cumsum(hit, axis=1) <= max_num_neighbors)
        rows, cols = np.nonzero(hit)
        counts[s:s + qq.shape[0]] = np.bincount(rows, minlength=qq.shape[0])
        idx_chunks.append(cols.astype(np.int64))
    splits = np.zeros(m + 1, dtype=np.int64)
    np.cumsum(counts, out=splits[1:])
    idx = np.concatenate(idx_chunks) if idx_chunks else np.zeros(0, dtype=np.int64)
    return {"neighbors_index": torch.from_numpy(idx), "neighbors_row_splits": torch.from_numpy(splits)}


def csr_transpose(nbrs, num_sources: int):
    """Source-sorted view of a CSR graph: for each source j the (edge id, row) pairs with
    neighbors_index[e]==j, ascending in e.  This is the structure agno backward gathers through."""
    idx = nbrs["neighbors_index"].numpy()
    rs = nbrs["neighbors_row_splits"].numpy()
    row = np.repeat(np.arange(rs.shape[0] - 1, dtype=np.int64), np.diff(rs))
    order = np.argsort(idx, kind="stable").astype(np.int64)
    cnt = np.bincount(idx, minlength=num_sources)
    splits = np.zeros(num_sources + 1, dtype=np.int64)
    np.cumsum(cnt, out=splits[1:])
    return {"neighbors_index": torch.from_numpy(row[order]), "neighbors_row_splits": torch.from_numpy(splits),
            "edge_perm": torch.from_numpy(order)}


# --------------------------------------------------------------------------------------------
# segment reductions (layers/utils/segment_csr.py:48-55 -> torch_scatter.segment_csr semantics)
# --------------------------------------------------------------------------------------------

def _seg_ids(splits):
    deg = splits[1:] - splits[:-1]
    return torch.repeat_interleave(torch.arange(deg.numel()), deg), deg


def segment_csr(src, splits, reduce: str):
    """Reduce dim -2 (or dim 0 for 1-D src) of `src` over CSR segments; empty segment -> 0;
    mean divides by max(deg, 1); max returns values only."""
    seg, deg = _seg_ids(splits)
    dim = 0 if src.dim() <= 2 else 1
    shape = list(src.shape)
    shape[dim] = deg.numel()
    out = torch.zeros(shape, dtype=src.dtype)
    if reduce in ("sum", "mean"):
        out = out.index_add(dim, seg, src)
        if reduce == "mean":
            view = [1] * src.dim()
            view[dim] = -1
            out = out / deg.clamp(min=1).to(src.dtype).view(view)
        return out
    if reduce == "max":
        view = [1] * src.dim()
        view[dim] = -1
        return out.scatter_reduce(dim, seg.view(view).expand_as(src), src, reduce="amax", include_self=False)
    raise ValueError(reduce)


# --------------------------------------------------------------------------------------------
# small layers (layers/mlp.py)
# --------------------------------------------------------------------------------------------

def linear_channel_mlp(p, prefix, x):
    """LinearChannelMLP.forward (mlp.py:329-337): Linear stack, exact-erf GELU between layers."""
    n = 0
    while f"{prefix}.fcs.{n}.weight" in p:
        n += 1
    for i in range(n):
        x = F.linear(x, p[f"{prefix}.fcs.{i}.weight"], p[f"{prefix}.fcs.{i}.bias"])
        if i < n - 1:
            x = F.gelu(x)
    return x


def channel_affine(p, prefix, x):
    """ChannelMLP(n_layers=1).forward on [B, rows, Cin] (mlp.py:283-305 with the permutes of
    magno.py:273-274): a single 1x1 Conv1d == affine map, no activation."""
    w = p[f"{prefix}.fcs.0.weight"][:, :, 0]
    return F.linear(x, w, p[f"{prefix}.fcs.0.bias"])


def node_pos_encode(x, freq: int = 4):
    """gemb.py:12-34."""
    freqs = torch.arange(1, freq + 1)
    phi = math.pi * (x + 1)
    z = freqs[None, :, None] * phi[:, None, :]
    z = torch.cat([z.sin(), z.cos()], dim=2)
    return z.reshape(x.shape[0], -1)


# --------------------------------------------------------------------------------------------
# A3. AGNO (layers/agno.py:148-273)
# --------------------------------------------------------------------------------------------

def agno_edge_weight(p, prefix, y, x, nbrs, use_attn=True, attention_type="cosine", coord_dim=None):
    """Batch-independent part of AGNO: w[e,:] = alpha_e * kappa_e (agno.py:212-250).
    kappa = channel_mlp(cat[y_j, x_m]) (agno.py:229, 242); alpha = segment-softmax of the cosine
    of the raw coordinate vectors (agno.py:218-224, 131-145) or of projected dot products
    (agno.py:214-217).  Without attention returns kappa (the reduce is then `mean`)."""
    idx, splits = nbrs["neighbors_index"], nbrs["neighbors_row_splits"]
    seg, _ = _seg_ids(splits)
    rep = y[idx]
    slf = x[seg]
    kappa = linear_channel_mlp(p, f"{prefix}.channel_mlp", torch.cat([rep, slf], dim=-1))
    if not use_attn:
        return kappa
    cd = coord_dim if coord_dim is not None else x.shape[1]
    qc, kc = slf[:, :cd], rep[:, :cd]
    if attention_type == "dot_product":
        qv = F.linear(qc, p[f"{prefix}.query_proj.weight"], p[f"{prefix}.query_proj.bias"])
        kv = F.linear(kc, p[f"{prefix}.key_proj.weight"], p[f"{prefix}.key_proj.bias"])
        s = (qv * kv).sum(-1) * (1.0 / 64 ** 0.5)
    else:
        s = (F.normalize(qc, p=2, dim=-1) * F.normalize(kc, p=2, dim=-1)).sum(-1)
    smax = segment_csr(s, splits, "max")[seg]
    ex = torch.exp(s - smax)
    den = segment_csr(ex, splits, "sum")[seg]
    return kappa * (ex / den).unsqueeze(-1)


def agno(p, prefix, y, nbrs, x, f_y, use_attn=True, attention_type="cosine", coord_dim=None):
    """AGNO.forward, transform_type='linear' (agno.py:245-271): out[b,m,:] = reduce_e w_e * f_y[b,idx_e,:],
    reduce = sum with attention, mean without."""
    w = agno_edge_weight(p, prefix, y, x, nbrs, use_attn, attention_type, coord_dim)
    msg = w.unsqueeze(0) * f_y[:, nbrs["neighbors_index"], :]
    return segment_csr(msg, nbrs["neighbors_row_splits"], "sum" if use_attn else "mean")


def agno_kernel_only(p, prefix, y, nbrs, x, use_attn=True, attention_type="cosine", coord_dim=None, weights=None):
    """AGNO.forward with f_y=None (transform (a), agno.py:229-271): out[m,:] = reduce_e w_e, reduce = sum with attention,
    mean without; with integration `weights` [n] (agno.py:252-262) every edge is scaled by weights[idx_e] and the reduce
    is a sum.  (The reference multiplies the [E] weight vector into the [E, C] kernel values without unsqueezing it, which
    only broadcasts for C == E or C == 1; this restates the documented per-edge scaling.)"""
    w = agno_edge_weight(p, prefix, y, x, nbrs, use_attn, attention_type, coord_dim)
    if weights is not None:
        return segment_csr(weights[nbrs["neighbors_index"]].unsqueeze(-1) * w, nbrs["neighbors_row_splits"], "sum")
    return segment_csr(w, nbrs["neighbors_row_splits"], "sum" if use_attn else "mean")


def agno_weighted(p, prefix, y, nbrs, x, f_y, weights, use_attn=True, attention_type="cosine", coord_dim=None):
    """AGNO.forward with f_y and integration weights (agno.py:245-262): sum_e weights[idx_e] w_e f_y[b, idx_e, :]."""
    w = agno_edge_weight(p, prefix, y, x, nbrs, use_attn, attention_type, coord_dim)
    msg = (weights[nbrs["neighbors_index"]].unsqueeze(-1) * w).unsqueeze(0) * f_y[:, nbrs["neighbors_index"], :]
    return segment_csr(msg, nbrs["neighbors_row_splits"], "sum")


# --------------------------------------------------------------------------------------------
# A4. statistical geometry embedding (layers/gemb.py:83-171, 230-233)
# --------------------------------------------------------------------------------------------

def geo_stat_features(input_geom, queries, nbrs):
    """gemb.py:103-169: [deg, mean dist, var dist, centroid - q (d), eig(cov) descending (d)],
    empty rows zeroed, then column standardisation (unbiased std, std<1e-6 -> 1)."""
    idx, splits = nbrs["neighbors_index"], nbrs["neighbors_row_splits"]
    seg, deg = _seg_ids(splits)
    m, d = queries.shape
    dt = queries.dtype
    n_i = deg.to(dt)
    has = deg > 0
    nc = n_i.clamp(min=1)
    pts = input_geom[idx]
    dist = torch.norm(pts - queries[seg], dim=1)
    d_avg = torch.zeros(m, dtype=dt).index_add(0, seg, dist) / nc
    e_x2 = torch.zeros(m, dtype=dt).index_add(0, seg, dist ** 2) / nc
    d_var = (e_x2 - d_avg ** 2).clamp(min=0.0)
    cen = torch.zeros(m, d, dtype=dt).index_add(0, seg, pts) / nc[:, None]
    delta = cen - queries
    ctr = pts - cen[seg]
    cov = torch.zeros(m, d, d, dtype=dt).index_add(0, seg, ctr.unsqueeze(2) * ctr.unsqueeze(1)) / nc.view(-1, 1, 1)
    pca = torch.zeros(m, d, dtype=dt)
    if has.any():
        pca[has] = torch.linalg.eigvalsh(cov[has]).flip(dims=[1])
    feat = torch.cat([n_i[:, None], d_avg[:, None], d_var[:, None], delta, pca], dim=1)
    feat[~has] = 0.0
    mean = feat.mean(dim=0, keepdim=True)
    std = feat.std(dim=0, keepdim=True)
    std[std < 1e-6] = 1.0
    return (feat - mean) / std


def geo_embed_pointnet(p, prefix, input_geom, queries, nbrs, pooling="max"):
    """GeometricEmbedding 'pointnet' (gemb.py:60-70, 173-228): MLP d->64->64 (ReLU, ReLU) on the neighbour coordinates
    centred on their query, scatter_max / scatter_mean over each row (empty rows 0), Linear 64->C + ReLU on the rows that
    have neighbours; the others stay zero."""
    idx, splits = nbrs["neighbors_index"], nbrs["neighbors_row_splits"]
    seg, deg = _seg_ids(splits)
    z = input_geom[idx] - queries[seg]
    h = F.relu(F.linear(z, p[f"{prefix}.pointnet_mlp.0.weight"], p[f"{prefix}.pointnet_mlp.0.bias"]))
    h = F.relu(F.linear(h, p[f"{prefix}.pointnet_mlp.2.weight"], p[f"{prefix}.pointnet_mlp.2.bias"]))
    pooled = segment_csr(h, splits, "max" if pooling == "max" else "mean")
    f = F.relu(F.linear(pooled, p[f"{prefix}.fc.0.weight"], p[f"{prefix}.fc.0.bias"]))
    return f * (deg > 0).to(f.dtype).unsqueeze(1)


def geo_embed(p, prefix, input_geom, queries, nbrs, method="statistical", pooling="max"):
    """GeometricEmbedding.forward 'statistical' (gemb.py:54-59, 230-233): ReLU(W2 ReLU(W1 f + b1) + b2)."""
    if method == "pointnet":
        return geo_embed_pointnet(p, prefix, input_geom, queries, nbrs, pooling)
    f = geo_stat_features(input_geom, queries, nbrs)
    h = F.relu(F.linear(f, p[f"{prefix}.mlp.0.weight"], p[f"{prefix}.mlp.0.bias"]))
    return F.relu(F.linear(h, p[f"{prefix}.mlp.2.weight"], p[f"{prefix}.mlp.2.bias"]))


# --------------------------------------------------------------------------------------------
# A2. MAGNO encoder / decoder (layers/magno.py:217-354, 552-692; vx loops :356-413, :694-751)
# --------------------------------------------------------------------------------------------

def _scale_weights(p, prefix, coords):
    h = F.relu(F.linear(coords, p[f"{prefix}.scale_weighting.0.weight"], p[f"{prefix}.scale_weighting.0.bias"]))
    return torch.softmax(F.linear(h, p[f"{prefix}.scale_weighting.2.weight"], p[f"{prefix}.scale_weighting.2.bias"]), dim=-1)


def _combine_scales(outs, weights):
    """magno.py:292-303 / 626-637."""
    if len(outs) == 1:
        return outs[0]
    if weights is not None:
        acc = torch.zeros_like(outs[0])
        for i, o in enumerate(outs):
            acc = acc + weights[:, i:i + 1].unsqueeze(0) * o
        return acc
    return torch.stack(outs, dim=0).mean(dim=0)


def _magno_one(p, prefix, mc, src_coord, dst_coord, feats, nbrs):
    """One scale, one graph: AGNO + geo-embed + recovery (magno.py:322-350 / 660-688)."""
    if mc.node_embedding:
        ys, xs = node_pos_encode(src_coord), node_pos_encode(dst_coord)
    else:
        ys, xs = src_coord, dst_coord
    out = agno(p, f"{prefix}.agno", ys, nbrs, xs, feats, mc.use_attention, mc.attention_type, ys.shape[1])
    if mc.use_geoembed:
        g = geo_embed(p, f"{prefix}.geoembed", src_coord, dst_coord, nbrs, mc.embedding_method, mc.pooling)
        out = torch.cat([out, g.unsqueeze(0).expand(out.shape[0], -1, -1)], dim=-1)
        out = channel_affine(p, f"{prefix}.recovery", out)
    return out


def _magno(p, prefix, mc, src_coord, dst_coord, feats, nbrs_in, src_is_batched):
    """Shared fx/vx driver.  `src_is_batched`: encoder vx has per-sample sources (x_coord[b]);
    decoder vx has per-sample destinations (query_coord[b])."""
    var = (src_coord if src_is_batched else dst_coord)
    vx = var.ndim == 3
    outs = []
    for s in range(len(mc.scales)):
        if not vx:
            outs.append(_magno_one(p, prefix, mc, src_coord, dst_coord, feats, nbrs_in[s]))
        else:
            per = []
            for b in range(var.shape[0]):
                sc = src_coord[b] if src_is_batched else src_coord
                dc = dst_coord if src_is_batched else dst_coord[b]
                per.append(_magno_one(p, prefix, mc, sc, dc, feats[b:b + 1], nbrs_in[b][s]))
            outs.append(torch.cat(per, dim=0))
    return outs


def build_graphs(mc, data, queries, inclusive=True, max_num_neighbors=None):
    """magno.py:174-215 (encoder: data=x_coord, queries=latent) / :510-550 (decoder: data=latent,
    queries=query_coord); radius per scale = radius*scale; vx = list[sample][scale]."""
    def one(dd, qq):
        return [radius_csr(dd, qq, mc.radius * s, inclusive, max_num_neighbors) for s in mc.scales]
    if data.ndim == 3:
        return [one(data[b], queries) for b in range(data.shape[0])]
    if queries.ndim == 3:
        return [one(data, queries[b]) for b in range(queries.shape[0])]
    return one(data, queries)


def magno_encoder(p, mc, x_coord, pndata, latent, nbrs=None, prefix="encoder"):
    if nbrs is None:
        nbrs = build_graphs(mc, x_coord, latent)
    f = channel_affine(p, f"{prefix}.lifting", pndata)
    outs = _magno(p, prefix, mc, x_coord, latent, f, nbrs, True)
    w = None
    if mc.use_scale_weights:
        w = _scale_weights(p, prefix, latent)
    return _combine_scales(outs, w)


def magno_decoder(p, mc, latent, rndata, query_coord, nbrs=None, prefix="decoder"):
    if nbrs is None:
        nbrs = build_graphs(mc, latent, query_coord)
    outs = _magno(p, prefix, mc, latent, query_coord, rndata, nbrs, False)
    w = None
    if mc.use_scale_weights:
        w = _scale_weights(p, prefix, query_coord if query_coord.ndim == 2 else query_coord[0])
    dec = _combine_scales(outs, w)
    return channel_affine(p, f"{prefix}.projection", dec)


# --------------------------------------------------------------------------------------------
# A5. processor (gaot.py:145-233, layers/attn.py)
# --------------------------------------------------------------------------------------------

def rmsnorm(x, w, eps):
    """attn.py:167-172 - the normalisation itself always runs in fp32 (`x.float()`), also for fp64 input."""
    xf = x.float()
    return (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)).type_as(x) * w


def conditioned_norm(p, prefix, c, x):
    """ConditionedNorm.forward (mlp.py:108-124) with its 2-"layer" MLPs (one Linear each, mlp.py:50-53):
    x * (1 + c * (Ws c + bs)) + c * (Wb c + bb), c [B, 1]."""
    scale = 1 + c * F.linear(c, p[f"{prefix}.mlp_scale.layers.0.weight"], p[f"{prefix}.mlp_scale.layers.0.bias"])
    bias = c * F.linear(c, p[f"{prefix}.mlp_bias.layers.0.weight"], p[f"{prefix}.mlp_bias.layers.0.bias"])
    return x * scale[:, None, :] + bias[:, None, :]


def rotary(t):
    """`RotaryEmbedding(dim=head_dim).rotate_queries_or_keys(t)` of the un-vendored rotary-embedding-torch wheel (attn.py:76,
    106-108) on [b, heads, seq, hd], library defaults: angle(pos, i) = pos / 10000^(2i / hd) shared by the pair (2i, 2i + 1),
    (x1, x2) -> (x1 cos - x2 sin, x2 cos + x1 sin)."""
    hd, T = t.shape[-1], t.shape[-2]
    inv = 1.0 / (10000 ** (torch.arange(0, hd, 2)[: hd // 2].float() / hd))
    ang = (torch.arange(T, dtype=inv.dtype)[:, None] * inv[None, :]).repeat_interleave(2, dim=-1).to(t.dtype)
    rot = torch.stack((-t[..., 1::2], t[..., 0::2]), dim=-1).reshape(t.shape)
    return t * ang.cos() + rot * ang.sin()


def attention(p, prefix, x, num_heads, num_kv_heads, condition=None, rope=False):
    """GroupQueryFlashAttention.forward (attn.py:78-119), non-causal, dropout 0; rope: positional_embedding='rope'."""
    if f"{prefix}.correction.mlp_scale.layers.0.weight" in p:        # attn.py:89-90
        x = conditioned_norm(p, f"{prefix}.correction", condition, x)
    b, t, _ = x.shape
    q = F.linear(x, p[f"{prefix}.q_proj.weight"])
    k = F.linear(x, p[f"{prefix}.k_proj.weight"])
    v = F.linear(x, p[f"{prefix}.v_proj.weight"])
    hd = q.shape[-1] // num_heads
    q = q.view(b, t, num_heads, hd).transpose(1, 2)
    k = k.view(b, t, num_kv_heads, hd).transpose(1, 2)
    v = v.view(b, t, num_kv_heads, hd).transpose(1, 2)
    if num_kv_heads != num_heads:
        k = k.repeat_interleave(num_heads // num_kv_heads, dim=1)
        v = v.repeat_interleave(num_heads // num_kv_heads, dim=1)
    if rope:                                              # attn.py:106-108
        q, k = rotary(q), rotary(k)
    o = F.scaled_dot_product_attention(q, k, v)           # attn.py:114 (non-causal, scale 1/sqrt(hd), dropout 0)
    o = o.transpose(1, 2).contiguous().view(b, t, -1)
    return F.linear(o, p[f"{prefix}.o_proj.weight"])


def ffn(p, prefix, x, condition=None):
    """FFN.forward SwiGLU (attn.py:150-156)."""
    y = F.linear(F.silu(F.linear(x, p[f"{prefix}.w1.weight"])) * F.linear(x, p[f"{prefix}.w3.weight"]),
                 p[f"{prefix}.w2.weight"])
    if f"{prefix}.correction.mlp_scale.layers.0.weight" in p:        # attn.py:153-154
        y = conditioned_norm(p, f"{prefix}.correction", condition, y)
    return y


def transformer_block(p, prefix, tc, x, skip=None, condition=None):
    """TransformerBlock.forward (attn.py:225-233) - note the non-standard residual: the second
    residual branch starts from the *normalised* h."""
    ac = tc.attn_config
    if skip is not None and f"{prefix}.skip_proj.weight" in p:
        x = F.linear(torch.cat([x, skip], dim=-1), p[f"{prefix}.skip_proj.weight"], p[f"{prefix}.skip_proj.bias"])
    h = rmsnorm(x, p[f"{prefix}.attn_norm.weight"], tc.norm_eps) if tc.use_attn_norm else x
    h = x + attention(p, f"{prefix}.attn", h, ac.num_heads, ac.num_kv_heads, condition,
                      rope=getattr(tc, "positional_embedding", "absolute") == "rope")
    h = rmsnorm(h, p[f"{prefix}.ffn_norm.weight"], tc.norm_eps) if tc.use_ffn_norm else h
    return h + ffn(p, f"{prefix}.ffn", h, condition)


def transformer(p, prefix, tc, x, condition=None):
    """Transformer.forward (attn.py:290-317): U-ViT with long skips."""
    if f"{prefix}.input_proj.weight" in p:
        x = F.linear(x, p[f"{prefix}.input_proj.weight"], p[f"{prefix}.input_proj.bias"])
    n_half = tc.num_layers // 2
    skips = []
    for i in range(n_half):
        x = transformer_block(p, f"{prefix}.encoder_layers.{i}", tc, x, None, condition)
        skips.append(x)
    if tc.num_layers % 2 == 1:
        x = transformer_block(p, f"{prefix}.middle_layer", tc, x, None, condition)
    for i in range(n_half):
        skip = skips.pop() if tc.use_long_range_skip else None
        x = transformer_block(p, f"{prefix}.decoder_layers.{i}", tc, x, skip, condition)
    if f"{prefix}.output_proj.weight" in p:
        x = F.linear(x, p[f"{prefix}.output_proj.weight"], p[f"{prefix}.output_proj.bias"])
    return x


def absolute_pos_emb(grid, embed_dim, dtype=torch.float32):
    """gaot.py:92-130: sin/cos of integer patch coordinates, base 10000."""
    pos = torch.stack(torch.meshgrid(*[torch.arange(g, dtype=torch.float32) for g in grid], indexing="ij"),
                      dim=-1).reshape(-1, len(grid))
    k = embed_dim // (2 * len(grid))
    inv = 1.0 / (10000 ** (torch.arange(k, dtype=torch.float32) / k))
    ang = pos[:, :, None] * inv[None, None, :]
    return torch.cat([torch.sin(ang), torch.cos(ang)], dim=-1).view(pos.shape[0], -1).to(dtype)


def patchify(x, latent_size, P):
    """gaot.py:182-185 (2-D) / :202-205 (3-D)."""
    b, _, c = x.shape
    if len(latent_size) == 2:
        H, W = latent_size
        x = x.view(b, H // P, P, W // P, P, c).permute(0, 1, 3, 2, 4, 5).contiguous()
        return x.view(b, (H // P) * (W // P), P * P * c)
    H, W, D = latent_size
    x = x.view(b, H // P, P, W // P, P, D // P, P, c).permute(0, 1, 3, 5, 2, 4, 6, 7).contiguous()
    return x.view(b, (H // P) * (W // P) * (D // P), P ** 3 * c)


def unpatchify(x, latent_size, P, c):
    """gaot.py:225-231."""
    b = x.shape[0]
    if len(latent_size) == 2:
        H, W = latent_size
        x = x.view(b, H // P, W // P, P, P, c).permute(0, 1, 3, 2, 4, 5).contiguous()
        return x.view(b, H * W, c)
    H, W, D = latent_size
    x = x.view(b, H // P, W // P, D // P, P, P, P, c).permute(0, 1, 4, 2, 5, 3, 6, 7).contiguous()
    return x.view(b, H * W * D, c)


def process(p, cfg, rndata, condition=None):
    """GAOT.process (gaot.py:145-233), positional_embedding='absolute'."""
    tc = cfg.args.transformer
    P = tc.patch_size
    c = rndata.shape[2]
    ls = tuple(cfg.latent_tokens_size)
    x = patchify(rndata, ls, P)
    x = F.linear(x, p["patch_linear.weight"], p["patch_linear.bias"])
    if getattr(tc, "positional_embedding", "absolute") == "absolute":    # gaot.py:212-218 ('rope' rotates q / k inside the attention)
        x = x + absolute_pos_emb([g // P for g in ls], x.shape[-1], x.dtype)
    x = transformer(p, "processor", tc, x, condition)
    return unpatchify(x, ls, P, c)


def gaot_forward(p, cfg, latent, xcoord, pndata, query_coord=None, encoder_nbrs=None, decoder_nbrs=None, condition=None):
    """GAOT.forward (gaot.py:248-305)."""
    mc = cfg.args.magno
    rn = magno_encoder(p, mc, xcoord, pndata, latent, encoder_nbrs)
    rn = process(p, cfg, rn, condition)
    if query_coord is None:
        query_coord = xcoord
    return magno_decoder(p, mc, latent, rn, query_coord, decoder_nbrs)


def train_step_loss(p, cfg, latent, xcoord, x, y, encoder_nbrs=None, decoder_nbrs=None, condition=None):
    """static_trainer.py:160-178 + base_trainer.py:71: MSE(mean) of the forward against y."""
    return F.mse_loss(gaot_forward(p, cfg, latent, xcoord, x, None, encoder_nbrs, decoder_nbrs, condition), y)


def adamw_step(params, grads, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=1e-2):
    """torch.optim.AdamW single step as driven by utils/optimizers.py:252-255 (torch defaults)."""
    for k in params:
        g = grads[k]
        params[k] = params[k] * (1 - lr * weight_decay)
        m[k] = beta1 * m[k] + (1 - beta1) * g
        v[k] = beta2 * v[k] + (1 - beta2) * g * g
        bc1 = 1 - beta1 ** step
        bc2 = 1 - beta2 ** step
        denom = (v[k].sqrt() / math.sqrt(bc2)) + eps
        params[k] = params[k] - (lr / bc1) * m[k] / denom
    return params, m, v


# --------------------------------------------------------------------------------------------
# rollout update rules (gaot.py:365-476)
# --------------------------------------------------------------------------------------------

def rollout_update(pred, u_in_norm, dt, stats, stepper_mode):
    """GAOT._process_autoregressive_prediction (gaot.py:436-476) and the state carry of gaot.py:425-434.
    Returns (pred_denorm, next_state): pred_denorm is what the reference appends to its output list (physical units);
    next_state = (pred_denorm - u_mean) / u_std is the normalised state fed to the next step (gaot.py:433)."""
    u_mean, u_std = stats["u"]["mean"], stats["u"]["std"]
    if stepper_mode == "output":
        u_next = pred * u_std + u_mean
    elif stepper_mode == "residual":
        u_next = (u_in_norm * u_std + u_mean) + (pred * stats["res"]["std"] + stats["res"]["mean"])
    elif stepper_mode == "time_der":
        u_next = (u_in_norm * u_std + u_mean) + dt * (pred * stats["der"]["std"] + stats["der"]["mean"])
    else:
        raise ValueError(f"Unsupported stepper_mode: {stepper_mode}")
    return u_next, (u_next - u_mean) / u_std


def autoregressive_predict(params, cfg, x_batch, time_indices, t_values, stats, stepper_mode, latent, coord,
                           use_conditional_norm=False):
    """GAOT.autoregressive_predict (gaot.py:307-434), fx mode: returns the DE-NORMALISED predictions
    [B, len(time_indices)-1, N, u_dim]; x_batch carries [u | c | ...] channels, c present iff "c" in stats."""
    B, N, _ = x_batch.shape
    u_dim = stats["u"]["mean"].shape[0]
    c_dim = stats["c"]["mean"].shape[0] if "c" in stats else 0
    c_feat = x_batch[..., u_dim:u_dim + c_dim] if c_dim > 0 else None
    cur = x_batch[..., :u_dim]
    outs = []
    for i in range(1, len(time_indices)):
        t0 = float(t_values[time_indices[i - 1]])
        dt = float(t_values[time_indices[i]]) - t0
        t0n = (t0 - stats["start_time"]["mean"]) / stats["start_time"]["std"]
        dtn = (dt - stats["time_diffs"]["mean"]) / stats["time_diffs"]["std"]
        feats = [cur] + ([c_feat] if c_feat is not None else [])
        feats += [torch.full((B, N, 1), t0n, dtype=x_batch.dtype), torch.full((B, N, 1), dtn, dtype=x_batch.dtype)]
        x_in = torch.cat(feats, dim=-1)
        with torch.no_grad():
            if use_conditional_norm:
                pred = gaot_forward(params, cfg, latent, coord, x_in[..., :-1], None, None, None, x_in[..., 0, -2:-1])
            else:
                pred = gaot_forward(params, cfg, latent, coord, x_in)
        phys, cur = rollout_update(pred, cur, dt, stats, stepper_mode)
        outs.append(phys)
    return torch.stack(outs, dim=1)


# --------------------------------------------------------------------------------------------
# config helpers (mirrors MAGNOConfig magno.py:26-69, TransformerConfig attn.py:19-38)
# --------------------------------------------------------------------------------------------

def make_config(latent_tokens_size, magno: Dict = None, transformer: Dict = None, attn: Dict = None):
    m = dict(coord_dim=2, radius=0.033, hidden_size=64, mlp_layers=3, lifting_channels=32, scales=[1.0],
             use_scale_weights=False, use_attention=True, attention_type="cosine", use_geoembed=True,
             embedding_method="statistical", pooling="max", transform_type="linear", sampling_strategy=None,
             max_neighbors=None, sample_ratio=None, node_embedding=False, neighbor_search_method="auto",
             use_torch_scatter=True, neighbor_strategy="radius", precompute_edges=False)
    m.update(magno or {})
    a = dict(num_heads=8, num_kv_heads=8, use_conditional_norm=False, cond_norm_hidden_size=4, atten_dropout=0.0)
    a.update(attn or {})
    t = dict(patch_size=8, hidden_size=256, use_attn_norm=True, use_ffn_norm=True, norm_eps=1e-6, num_layers=3,
             positional_embedding="absolute", use_long_range_skip=True, ffn_multiplier=4)
    t.update(transformer or {})
    t["attn_config"] = SimpleNamespace(**a)
    return SimpleNamespace(latent_tokens_size=tuple(latent_tokens_size),
                           args=SimpleNamespace(magno=SimpleNamespace(**m), transformer=SimpleNamespace(**t)))
