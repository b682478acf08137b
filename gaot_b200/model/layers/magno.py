"""MAGNO encoder / decoder (reference src/model/layers/magno.py:26-751) on the sm_100a kernels.

fx mode (coords [N,d] shared by the batch): one graph, the batch is the inner loop of every kernel.
vx mode (coords [B,N,d]): the per-sample graphs are merged block-diagonally (ops.merge_graphs) so the
whole batch is still ONE launch per operator instead of the reference's Python loop over samples
(magno.py:361-411, 699-749).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import torch
from torch import nn

from ... import ops
from .agno import AGNO
from .gemb import GeometricEmbedding, node_pos_encode
from .mlp import ChannelMLP
from .utils.edge_drop import apply_edge_drop_csr
from .utils.neighbor_search import NeighborSearch


@dataclass
class MAGNOConfig:
    """Field-for-field the reference MAGNOConfig (magno.py:26-69)."""
    coord_dim: int = 2
    radius: float = 0.033
    hidden_size: int = 64
    mlp_layers: int = 3
    lifting_channels: int = 32
    scales: List[float] = field(default_factory=lambda: [1.0])
    use_scale_weights: bool = False
    use_attention: bool = True
    attention_type: str = "cosine"
    use_geoembed: bool = True
    embedding_method: str = "statistical"
    pooling: str = "max"
    transform_type: str = "linear"
    sampling_strategy: Optional[str] = None
    max_neighbors: Optional[int] = None
    sample_ratio: Optional[float] = None
    node_embedding: bool = False
    neighbor_search_method: str = "auto"
    use_torch_scatter: bool = True
    neighbor_strategy: str = "radius"
    precompute_edges: bool = False

    def __post_init__(self):
        if self.coord_dim not in [2, 3]:
            raise ValueError(f"coord_dim must be 2 or 3, got {self.coord_dim}")
        if self.sampling_strategy == "ratio" and (self.sample_ratio is None or not 0 < self.sample_ratio <= 1):
            raise ValueError("sample_ratio must be in (0, 1] when using 'ratio' sampling")
        if self.sampling_strategy == "max_neighbors" and (self.max_neighbors is None or self.max_neighbors <= 0):
            raise ValueError("max_neighbors must be > 0 when using 'max_neighbors' sampling")


class _MAGNOBase(nn.Module):
    """What encoder and decoder share: graph handling, one-scale operator, scale combination."""

    def _setup(self, config, agno_out: int):
        self.config = config
        self.coord_dim = config.coord_dim
        self.scales = config.scales
        self.use_scale_weights = config.use_scale_weights
        self.precompute_edges = config.precompute_edges
        self.use_geoembed = config.use_geoembed
        self.node_embedding = config.node_embedding
        self.nb_search = NeighborSearch(method=config.neighbor_search_method)
        self.neighbor_cache = {}
        self._merged_cache = {}
        self.fuse_affine = True          # fold the affine lifting / recovery+projection into the graph kernels
        self.sampling_strategy = config.sampling_strategy
        self.max_neighbors = config.max_neighbors
        self.sample_ratio = config.sample_ratio
        kcd = self._compute_kernel_coord_dim()
        kin = kcd * 2
        if config.transform_type in ["nonlinear", "nonlinear_kernelonly"]:
            raise NotImplementedError("nonlinear transform types are shape-broken in the reference (magno.py:112-116 vs :273-274)")
        self.agno = AGNO(channel_mlp_layers=[kin] + [config.hidden_size] * config.mlp_layers + [agno_out],
                         transform_type=config.transform_type, use_attn=config.use_attention,
                         attention_type=config.attention_type, coord_dim=kcd, use_torch_scatter=config.use_torch_scatter)
        return kcd

    def _finish_setup(self, config, channels: int, kcd: int):
        if self.use_geoembed:
            self.geoembed = GeometricEmbedding(input_dim=self.coord_dim, output_dim=channels,
                                               method=config.embedding_method, pooling=config.pooling)
            self.recovery = ChannelMLP(in_channels=2 * channels, out_channels=channels, n_layers=1)
        if self.use_scale_weights:
            self.scale_weighting = nn.Sequential(nn.Linear(kcd, config.hidden_size // 4), nn.ReLU(),
                                                 nn.Linear(config.hidden_size // 4, len(self.scales)))
            self.scale_weight_activation = nn.Softmax(dim=-1)

    def _compute_kernel_coord_dim(self) -> int:
        return self.coord_dim * 4 * 2 if self.node_embedding else self.coord_dim

    # ---- graphs ---------------------------------------------------------------------------------
    def _search(self, data, queries):
        """list[scale] (fx) or list[sample][scale] (vx) of CSR dicts (magno.py:174-215 / 510-550)."""
        key = (data.data_ptr(), data._version, tuple(data.shape), queries.data_ptr(), queries._version, tuple(queries.shape))
        hit = self.neighbor_cache.get("graph")
        if hit is not None and hit[0] == key:
            return hit[1]
        with torch.no_grad():
            per_scale = [self.nb_search(data, queries, self.config.radius * s) for s in self.scales]
        if data.ndim == 3 or queries.ndim == 3:
            nb = len(per_scale[0])
            graphs = [[per_scale[s][b] for s in range(len(self.scales))] for b in range(nb)]
        else:
            graphs = per_scale
        self.neighbor_cache = {"graph": (key, graphs, data, queries)}   # keeps the keyed tensors alive
        return graphs

    def _merged_plan(self, graphs_b: List[dict], n_src_each: int) -> ops.GraphPlan:
        key = tuple(id(g) for g in graphs_b)
        hit = self._merged_cache.get(key)
        if hit is None:
            if len(self._merged_cache) >= 8:
                self._merged_cache.clear()
            hit = (ops.merge_graphs(graphs_b, n_src_each), graphs_b)
            self._merged_cache[key] = hit
        return hit[0]

    # ---- one scale --------------------------------------------------------------------------------
    def _edge_weights(self, plan, src_coord, dst_coord):
        ys, xs = (node_pos_encode(src_coord), node_pos_encode(dst_coord)) if self.node_embedding else (src_coord, dst_coord)
        return self.agno.edge_weights(ys, xs, plan)

    # ---- destination rows sharded over the data-parallel ranks (ops.enable_edge_sharding) --------------------------------
    @staticmethod
    def _row_sharding(plan, n_rows: int):
        """(world, rank, group) when the MAGNO operator on this graph runs row-sharded: every rank owns 1 / world of the
        destination rows for ALL samples of the global batch (edge-weight MLP, gather / reduce and their backward on that
        slice only), samples are exchanged with one all-gather in and one all-to-all out.  Otherwise world = 1."""
        world, rank, grp = ops.edge_shard_world()
        if world > 1 and plan.E >= ops.SHARD_EDGES_MIN and n_rows % world == 0:
            return world, rank, grp
        return 1, 0, None

    def _one_scale(self, plan, src_coord, dst_coord, feats, groups: int, lift=None):
        """AGNO + geo-embedding + recovery on one graph (magno.py:322-350 / 660-688).
        src/dst coords are [n_src,d] / [m,d] matching the plan; feats [Bf, n_src, C] - or, with
        lift=(Wl, bl), the raw inputs [Bf, n_src, Cin] whose affine lifting is fused into the gather."""
        world, rank, grp = self._row_sharding(plan, plan.m) if groups == 1 else (1, 0, None)
        if world > 1:
            # my slice of the latent rows for the WHOLE global batch, then samples <-> rows exchange; the edge weights of the
            # other slices are never computed on this rank
            sl = plan.row_slice(rank, world)
            w = self._edge_weights(sl, src_coord, dst_coord[sl.row_lo:sl.row_hi])
            f_all = ops.all_gather_batch(feats, grp)                      # [B_global, n_src, Cin | C]
            out_s = ops.lift_aggregate(f_all, lift[0], lift[1], w, sl) if lift is not None else ops.agno_aggregate(f_all, w, sl)
            out = ops.rows_to_samples(out_s, grp)                         # [B_local, m, C]
        else:
            w = self._edge_weights(plan, src_coord, dst_coord)
            if lift is not None:
                out = ops.lift_aggregate(feats, lift[0], lift[1], w, plan)   # [Bf, m, C], lifted features never materialised
            else:
                out = ops.agno_aggregate(feats, w, plan)                     # [Bf, m, C]
        if self.use_geoembed:
            C = out.shape[-1]
            feat = self.geoembed.features(src_coord, dst_coord, plan, groups, cache=(groups == 1))
            g = self.geoembed.embed(feat, tc=self._embed_on_tc(groups, feat))   # [m, C], batch independent
            W = self.recovery.fcs[0].weight[:, :, 0]                     # [C, 2C]: cat[agno, geo] (magno.py:346-350)
            gterm = ops.linear(g, W[:, C:], self.recovery.fcs[0].bias, tc=self._embed_on_tc(groups, feat))   # batch-independent half of the recovery
            # [B*m, C] x [C, C]: the one big dense layer of MAGNO; tensor cores in bf16 mode, FFMA in fp32 mode
            out = ops.linear(out.view(-1, C), W[:, :C], None, add=gterm, tc=True).view(out.shape)
        return out

    def _combine(self, outs, weight_coords):
        if len(outs) == 1:
            return outs[0]
        if self.use_scale_weights:
            h = ops.linear(weight_coords, self.scale_weighting[0].weight, self.scale_weighting[0].bias, act=ops.ACT_RELU)
            sw = torch.softmax(ops.linear(h, self.scale_weighting[2].weight, self.scale_weighting[2].bias), dim=-1)
            acc = None
            for i, o in enumerate(outs):
                t = o * sw[:, i:i + 1].unsqueeze(0)
                acc = t if acc is None else acc + t
            return acc
        return torch.stack(outs, 0).mean(dim=0)

    EMBED_TC_MIN_ROWS = 1 << 16

    @classmethod
    def _embed_on_tc(cls, groups: int, feat) -> bool:
        """vx mode embeds B * rows points per step: from 65 536 rows on, the K = 64 layers of the geometry embedding go to
        the tensor cores in bf16 mode; fx mode (a once-per-graph-sized side computation) and small batches stay fp32."""
        rows = feat.shape[0] if torch.is_tensor(feat) else feat[1].shape[0]
        return groups > 1 and rows >= cls.EMBED_TC_MIN_ROWS

    def _drop(self, nbrs):
        return apply_edge_drop_csr(nbrs, self.sampling_strategy, self.max_neighbors, self.sample_ratio, self.training)


class MAGNOEncoder(_MAGNOBase):
    """magno.py:74-413.  forward(x_coord [N,d]|[B,N,d], pndata [B,N,Cin], latent [M,d]) -> [B,M,C]."""

    def __init__(self, in_channels: int, out_channels: int, config: MAGNOConfig):
        super().__init__()
        kcd = self._setup(config, out_channels)
        self.lifting = ChannelMLP(in_channels=in_channels, hidden_channels=config.hidden_size,
                                  out_channels=out_channels, n_layers=1)
        self._finish_setup(config, out_channels, kcd)

    def forward(self, x_coord, pndata, latent_tokens_coord, encoder_nbrs=None):
        if x_coord.ndim not in (2, 3):
            raise ValueError(f"x_coord must be 2D [n_nodes, coord_dim] or 3D [batch_size, n_nodes, coord_dim], got {x_coord.ndim}D")
        vx = x_coord.ndim == 3
        if x_coord.shape[-1] != self.coord_dim or latent_tokens_coord.shape[-1] != self.coord_dim:
            raise ValueError(f"Expected coord_dim={self.coord_dim}, got {x_coord.shape[-1]} / {latent_tokens_coord.shape[-1]}")
        if pndata.shape[1] != x_coord.shape[-2]:
            raise ValueError(f"pndata has {pndata.shape[1]} nodes but x_coord has {x_coord.shape[-2]}")
        if vx and pndata.shape[0] != x_coord.shape[0]:
            raise ValueError("batch size mismatch between pndata and x_coord")
        if self.precompute_edges:
            if encoder_nbrs is None:
                raise ValueError("encoder_nbrs is required when precompute_edges is True")
        else:
            encoder_nbrs = self._search(x_coord, latent_tokens_coord)
        B, N = pndata.shape[0], pndata.shape[1]
        M = latent_tokens_coord.shape[0]
        Wl, bl = self.lifting.fcs[0].weight[:, :, 0], self.lifting.fcs[0].bias
        C, Cin = Wl.shape
        # the fused lift+gather kernel produces no gradient for its raw inputs (they are data at every BASELINE config);
        # when something upstream does need d/d pndata the unfused lifting + aggregate path propagates it
        fuse = (self.fuse_affine and Cin <= ops.FUSE_MAX_CIN and C % 4 == 0 and C <= ops.FUSE_MAX_C
                and not (pndata.requires_grad and torch.is_grad_enabled()))
        lift = (Wl, bl) if fuse else None
        f = pndata if fuse else self.lifting(pndata)                      # [B,N,C]  (magno.py:273-274)
        Cf = f.shape[-1]
        outs = []
        for s in range(len(self.scales)):
            if not vx:
                nb = self._drop(encoder_nbrs[s])
                plan = ops.graph_plan(nb, N)
                outs.append(self._one_scale(plan, x_coord, latent_tokens_coord, f, 1, lift))
            else:
                plan = self._merged_plan([self._drop(encoder_nbrs[b][s]) for b in range(B)], N)
                o = self._one_scale(plan, x_coord.reshape(B * N, -1), latent_tokens_coord.repeat(B, 1),
                                    f.reshape(1, B * N, Cf), B, lift)
                outs.append(o.view(B, M, C))
        return self._combine(outs, latent_tokens_coord)


class MAGNODecoder(_MAGNOBase):
    """magno.py:418-751.  forward(latent [M,d], rndata [B,M,C], query_coord [Nq,d]|[B,Nq,d]) -> [B,Nq,Cout]."""

    def __init__(self, in_channels: int, out_channels: int, config: MAGNOConfig):
        super().__init__()
        kcd = self._setup(config, in_channels)
        self.projection = ChannelMLP(in_channels=in_channels, hidden_channels=config.hidden_size,
                                     out_channels=out_channels, n_layers=1)
        self._finish_setup(config, in_channels, kcd)

    def forward(self, latent_tokens_coord, rndata, query_coord, decoder_nbrs=None):
        if query_coord.ndim not in (2, 3):
            raise ValueError(f"query_coord must be 2D or 3D, got {query_coord.ndim}D")
        vx = query_coord.ndim == 3
        if query_coord.shape[-1] != self.coord_dim or latent_tokens_coord.shape[-1] != self.coord_dim:
            raise ValueError(f"Expected coord_dim={self.coord_dim}")
        if rndata.shape[1] != latent_tokens_coord.shape[0]:
            raise ValueError(f"rndata has {rndata.shape[1]} tokens but latent_tokens_coord has {latent_tokens_coord.shape[0]}")
        if self.precompute_edges:
            if decoder_nbrs is None:
                raise ValueError("decoder_nbrs is required when precompute_edges is True")
        else:
            decoder_nbrs = self._search(latent_tokens_coord, query_coord)
        B, M, C = rndata.shape
        Nq = query_coord.shape[-2]
        Wp, bp = self.projection.fcs[0].weight[:, :, 0], self.projection.fcs[0].bias
        if (self.fuse_affine and len(self.scales) == 1 and Wp.shape[0] <= ops.FUSE_MAX_COUT and C % 4 == 0
                and C <= ops.FUSE_MAX_C):
            return self._forward_fused(latent_tokens_coord, rndata, query_coord, decoder_nbrs, vx, Wp, bp)
        outs = []
        for s in range(len(self.scales)):
            if not vx:
                nb = self._drop(decoder_nbrs[s])
                plan = ops.graph_plan(nb, M)
                outs.append(self._one_scale(plan, latent_tokens_coord, query_coord, rndata, 1))
            else:
                plan = self._merged_plan([self._drop(decoder_nbrs[b][s]) for b in range(B)], M)
                o = self._one_scale(plan, latent_tokens_coord.repeat(B, 1), query_coord.reshape(B * Nq, -1),
                                    rndata.reshape(1, B * M, C), B)
                outs.append(o.view(B, Nq, C))
        dec = self._combine(outs, query_coord if not vx else query_coord[0])
        return self.projection(dec)                                       # [B,Nq,Cout]  (magno.py:640-641)

    def fusable(self, C: int) -> bool:
        Wp = self.projection.fcs[0].weight
        return (self.fuse_affine and len(self.scales) == 1 and Wp.shape[0] <= ops.FUSE_MAX_COUT and C % 4 == 0
                and C <= ops.FUSE_MAX_C)

    def prepare_fused(self, latent, query_coord, decoder_nbrs, B: int, C: int, resolved: bool = False):
        """Batch-independent half of the fused decoder: graph plan, edge weights (kernel MLP + attention), geometry term and
        the folded read-out weights.  It depends on coordinates and parameters only, so GAOT.forward runs it on a side
        stream while the processor is busy (its many small latency-bound kernels then cost nothing); the backward of these
        ops runs on the same side stream under the processor's backward."""
        if query_coord.ndim not in (2, 3):
            raise ValueError(f"query_coord must be 2D or 3D, got {query_coord.ndim}D")
        vx = query_coord.ndim == 3
        if query_coord.shape[-1] != self.coord_dim or latent.shape[-1] != self.coord_dim:
            raise ValueError(f"Expected coord_dim={self.coord_dim}")
        if not resolved:
            if self.precompute_edges:
                if decoder_nbrs is None:
                    raise ValueError("decoder_nbrs is required when precompute_edges is True")
            else:
                decoder_nbrs = self._search(latent, query_coord)
        M = latent.shape[0]
        Nq = query_coord.shape[-2]
        Wp, bp = self.projection.fcs[0].weight[:, :, 0], self.projection.fcs[0].bias
        if not vx:
            plan = ops.graph_plan(self._drop(decoder_nbrs[0]), M)
            src, dst, groups = latent, query_coord, 1
        else:
            plan = self._merged_plan([self._drop(decoder_nbrs[b][0]) for b in range(B)], M)
            src, dst, groups = latent.repeat(B, 1), query_coord.reshape(B * Nq, -1), B
        world, rank, grp = self._row_sharding(plan, Nq) if not vx else (1, 0, None)
        full_plan = plan
        if world > 1:
            # row-sharded: this rank prepares its slice of the query rows only (edge weights, geometry term, read-out
            # weights); apply_fused runs that slice for the whole global batch and exchanges samples <-> rows
            plan = full_plan.row_slice(rank, world)
            dst = dst[plan.row_lo:plan.row_hi]
        w = self._edge_weights(plan, src, dst)                            # [E, C]
        if self.use_geoembed:
            Wrec = self.recovery.fcs[0].weight[:, :, 0]                   # [C, 2C]
            if world > 1:
                # the statistics are standardised over ALL rows (gemb.py:160-166): computed on the full graph (cached), sliced
                feat = self.geoembed.features(src, query_coord, full_plan, 1, cache=True)
                if torch.is_tensor(feat):
                    feat = feat[plan.row_lo:plan.row_hi]
                else:
                    raise NotImplementedError("row sharding with the pointnet embedding")
            else:
                feat = self.geoembed.features(src, dst, plan, groups, cache=(groups == 1))
            g = self.geoembed.embed(feat, tc=self._embed_on_tc(groups, feat))   # [rows, C]
            # projection o recovery folded on the parameter side: const = (Wp Wg) g + (Wp b_rec + bp); the [rows, C]
            # recovery output of the geometry half is never formed (rows = B * Nq in vx mode)
            Wgeo = ops.linear(Wp, Wrec[:, C:].t())                        # Wp @ Wg   [Cout, C]
            bgeo = ops.linear(self.recovery.fcs[0].bias.view(1, -1), Wp, bp).view(-1)
            const = ops.linear(g, Wgeo, bgeo)                             # [rows, Cout]
            Wfold = ops.linear(Wp, Wrec[:, :C].t())                       # Wp @ Wa   [Cout, C]
        else:
            const, Wfold = bp.view(1, -1), Wp
        u = Wfold.unsqueeze(1) * w.unsqueeze(0)                           # [Cout, E, C]
        return dict(plan=plan, u=u, const=const, vx=vx, Nq=Nq, shard=(world, grp), n_src=M)

    def apply_fused(self, rndata, prep):
        B, M, C = rndata.shape
        if rndata.shape[1] != prep["plan"].n_src // (B if prep["vx"] else 1):   # (a row slice keeps the full source numbering)
            raise ValueError(f"rndata has {rndata.shape[1]} tokens but latent_tokens_coord has {prep['plan'].n_src // (B if prep['vx'] else 1)}")
        world, grp = prep.get("shard", (1, None))
        if world > 1:
            r_all = ops.all_gather_batch(rndata, grp)                     # [B_global, M, C]: the latent of every sample
            out_s = ops.agno_dot(r_all, prep["u"], prep["plan"]) + prep["const"].unsqueeze(0)   # my query rows, all samples
            return ops.rows_to_samples(out_s, grp)                        # [B_local, Nq, Cout]
        feats = rndata.reshape(1, B * M, C) if prep["vx"] else rndata
        out = ops.agno_dot(feats, prep["u"], prep["plan"])               # [Bf, rows, Cout]
        return (out + prep["const"].unsqueeze(0)).view(B, prep["Nq"], -1)

    def _forward_fused(self, latent, rndata, query_coord, decoder_nbrs, vx, Wp, bp):
        """Single-scale decoder with recovery and projection (both affine, no activation between them:
        magno.py:684-688, 640-641) folded into the aggregate:
            out[b,n,:] = (Wp Wa) A[b,n,:] + Wp (Wg G[n,:] + b_rec) + bp,   A = AGNO output
        so the kernel reduces <u[o,e,:], r[b,j,:]> with u = (Wp Wa)[o,:] * w[e,:] and writes Cout scalars per
        point; the [B,N,C] decoder activations are never formed."""
        B, M, C = rndata.shape
        prep = self.prepare_fused(latent, query_coord, decoder_nbrs, B, C, resolved=True)
        return self.apply_fused(rndata, prep)
