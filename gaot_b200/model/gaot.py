"""GAOT top level (reference src/model/gaot.py:11-476): MAGNO encoder -> patchify + U-ViT -> MAGNO decoder,
same constructor, forward signature and state-dict keys; all arithmetic in libgaot_b200.so."""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from .. import ops
from .layers.attn import Transformer
from .layers.magno import MAGNODecoder, MAGNOEncoder


class GAOT(nn.Module):
    def __init__(self, input_size: int, output_size: int, config=None):
        super().__init__()
        coord_dim = config.args.magno.coord_dim
        if coord_dim not in [2, 3]:
            raise ValueError(f"coord_dim must be 2 or 3, got {coord_dim}")
        self.input_size, self.output_size, self.coord_dim = input_size, output_size, coord_dim
        self.node_latent_size = config.args.magno.lifting_channels
        self.patch_size = config.args.transformer.patch_size
        lts = tuple(config.latent_tokens_size)
        if len(lts) != coord_dim:
            raise ValueError(f"For {coord_dim}D, latent_tokens_size must have {coord_dim} dimensions, got {len(lts)}")
        self.H, self.W = lts[0], lts[1]
        self.D = lts[2] if coord_dim == 3 else None
        self.encoder = MAGNOEncoder(in_channels=input_size, out_channels=self.node_latent_size, config=config.args.magno)
        pv = self.patch_size ** coord_dim
        self.patch_linear = nn.Linear(pv * self.node_latent_size, pv * self.node_latent_size)
        self.positional_embedding_name = config.args.transformer.positional_embedding
        self.positions = self._get_patch_positions()
        self.processor = Transformer(input_size=self.node_latent_size * pv, output_size=self.node_latent_size * pv,
                                     config=config.args.transformer)
        self.decoder = MAGNODecoder(in_channels=self.node_latent_size, out_channels=output_size, config=config.args.magno)
        self.overlap_decoder_prep = True        # run the decoder's batch-independent work on a side stream (CUDA only)
        self._side_stream = None
        self._pos_emb = None
        self._rollout_graph = None              # captured rollout forward (autoregressive_predict(use_cuda_graph=True))

    def _grid(self):
        return (self.H, self.W) if self.coord_dim == 2 else (self.H, self.W, self.D)

    def _get_patch_positions(self):
        axes = [torch.arange(g // self.patch_size, dtype=torch.float32) for g in self._grid()]
        return torch.stack(torch.meshgrid(*axes, indexing="ij"), dim=-1).reshape(-1, self.coord_dim)

    def _compute_absolute_embeddings(self, positions, embed_dim):
        """gaot.py:119-130 - a constant table, built once per device."""
        nd = positions.size(1)
        k = embed_dim // (2 * nd)
        inv = 1.0 / (10000 ** (torch.arange(k, dtype=torch.float32, device=positions.device) / k))
        ang = positions[:, :, None] * inv[None, None, :]
        return torch.cat([torch.sin(ang), torch.cos(ang)], dim=-1).view(positions.size(0), -1)

    def encode(self, x_coord, pndata, latent_tokens_coord, encoder_nbrs):
        return self.encoder(x_coord=x_coord, pndata=pndata, latent_tokens_coord=latent_tokens_coord,
                            encoder_nbrs=encoder_nbrs)

    def process(self, rndata, condition: Optional[float] = None):
        """gaot.py:145-233."""
        B, n_tok, C = rndata.shape
        grid, P = self._grid(), self.patch_size
        n_expected = 1
        for g in grid:
            n_expected *= g
            if g % P != 0:
                raise ValueError(f"latent grid {grid} must be divisible by patch size P={P}")
        if n_tok != n_expected:
            raise ValueError(f"rndata has {n_tok} tokens, expected {n_expected} for latent grid {grid}")
        x = ops.patchify(rndata, grid, P)                                 # [B,T,P^d*C]
        T, Dp = x.shape[1], x.shape[2]
        add = None
        if self.positional_embedding_name == "absolute":
            if self._pos_emb is None or self._pos_emb.device != x.device:
                pe = self._compute_absolute_embeddings(self.positions.to(x.device), Dp)
                if pe.shape[1] != Dp:
                    raise ValueError(f"absolute positional embedding width {pe.shape[1]} != patch dim {Dp}")
                self._pos_emb = pe.contiguous()
            add = self._pos_emb
        relative_positions = self.positions if self.positional_embedding_name == "rope" else None       # gaot.py:217-218
        x = ops.linear(x.view(B * T, Dp), self.patch_linear.weight, self.patch_linear.bias, add=add, tc=True).view(B, T, Dp)
        x = self.processor(x, condition=condition, relative_positions=relative_positions)
        return ops.unpatchify(x, grid, P, C)

    def decode(self, latent_tokens_coord, rndata, query_coord, decoder_nbrs):
        return self.decoder(latent_tokens_coord=latent_tokens_coord, rndata=rndata, query_coord=query_coord,
                            decoder_nbrs=decoder_nbrs)

    def forward(self, latent_tokens_coord, xcoord, pndata, query_coord=None, encoder_nbrs=None, decoder_nbrs=None,
                condition: Optional[float] = None):
        if query_coord is None:
            query_coord = xcoord
        ops.clear_prepared()                               # prepared bf16 weights never outlive the forward that made them
        C = self.decoder.recovery.fcs[0].weight.shape[0] if self.decoder.use_geoembed else self.decoder.projection.fcs[0].weight.shape[1]
        if self.overlap_decoder_prep and pndata.is_cuda and self.decoder.fusable(C):
            # The decoder's batch-independent work (edge MLP, attention weights, geometry term: ~30 small latency-bound
            # kernels forward + backward) depends on coordinates and parameters only: fork it onto a side stream so it runs
            # under the processor (in a captured CUDA graph the two branches replay concurrently; autograd runs the
            # backward of these ops on the same side stream, i.e. under the processor's backward).
            cur = torch.cuda.current_stream()
            if self._side_stream is None:
                self._side_stream = torch.cuda.Stream()
            side = self._side_stream
            side.wait_stream(cur)
            wready = None
            with torch.cuda.stream(side):
                if ops.get_precision() in ("bf16", "fp32"):
                    # tensor-core operands (bf16 copies, or fp16 hi / lo planes in the parity mode) and concatenations of the
                    # processor's weights: ~40 tiny kernels that would otherwise sit between the processor's GEMMs; here
                    # they run under the encoder
                    ops.clear_prepared()
                    wts = self.processor.prep_weights() + ops.prepare_weight(self.patch_linear.weight)
                    wready = torch.cuda.Event()
                    wready.record(side)
                prep = self.decoder.prepare_fused(latent_tokens_coord, query_coord, decoder_nbrs, pndata.shape[0], C)
            rndata = self.encode(xcoord, pndata, latent_tokens_coord, encoder_nbrs)
            if wready is not None:
                cur.wait_event(wready)
                for t in wts:
                    t.record_stream(cur)
            rndata = self.process(rndata, condition=condition)
            ops.clear_prepared()
            cur.wait_stream(side)
            for t in (prep["u"], prep["const"]):
                t.record_stream(cur)                       # produced on the side stream, consumed on this one
            return self.decoder.apply_fused(rndata, prep)
        rndata = self.encode(xcoord, pndata, latent_tokens_coord, encoder_nbrs)
        rndata = self.process(rndata, condition=condition)
        return self.decode(latent_tokens_coord, rndata, query_coord, decoder_nbrs)

    @torch.no_grad()
    def autoregressive_predict(self, x_batch, time_indices, t_values, stats, stepper_mode="output",
                               latent_tokens_coord=None, fixed_coord=None, encoder_nbrs=None, decoder_nbrs=None,
                               use_conditional_norm=False, use_cuda_graph=False):
        """gaot.py:307-434, same arguments in the same order (use_cuda_graph is the one addition).
        x_batch [B,N,>=u_dim+c_dim]: normalised initial state in the first u_dim = len(stats["u"]["mean"]) channels, static
        conditioning channels next when stats has a "c" entry (gaot.py:349-358); anything after them (the time channels of a
        dataset sample) is ignored, as in the reference.  Returns the DE-NORMALISED predictions
        [B, len(time_indices)-1, N, u_dim] (gaot.py:425-434); only the carried state is re-normalised.
        stats = {"u": {mean,std}, "res": {...}, "der": {...}, "c": {...}?, "start_time": {...}, "time_diffs": {...}} as the
        reference's sequential trainer supplies them.  With use_conditional_norm the last input channel (the time difference) is
        not fed to the encoder and the conditioning value is x_input[..., 0, -2:-1], exactly as gaot.py:392-408 does."""
        if stepper_mode not in ("output", "residual", "time_der"):
            raise ValueError(f"Unsupported stepper_mode: {stepper_mode}")
        dev = x_batch.device
        B, N, _ = x_batch.shape
        U = int(torch.as_tensor(stats["u"]["mean"]).numel())
        nc = int(torch.as_tensor(stats["c"]["mean"]).numel()) if "c" in stats else 0
        cur = x_batch[..., :U].contiguous().float()
        outs = []
        mode = {"output": 0, "residual": 1, "time_der": 2}[stepper_mode]
        t = lambda v: torch.as_tensor(v, dtype=torch.float32, device=dev).reshape(-1).expand(U).contiguous()  # noqa: E731
        um, us = t(stats["u"]["mean"]), t(stats["u"]["std"])
        skey = {1: "res", 2: "der"}.get(mode)
        sm, ss = (t(stats[skey]["mean"]), t(stats[skey]["std"])) if skey else (um, us)
        # static step input [u | c | t0 | dt]; with use_cuda_graph the buffer, the captured forward and its output live on the
        # model and are reused by every later rollout over the same mesh and batch shape (capture costs ~30 forwards)
        key = (B, N, U, nc, bool(use_conditional_norm), ops.get_precision(),
               None if latent_tokens_coord is None else latent_tokens_coord.data_ptr(),
               None if fixed_coord is None else fixed_coord.data_ptr(), id(encoder_nbrs), id(decoder_nbrs))
        cached = self._rollout_graph if (use_cuda_graph and dev.type == "cuda" and self._rollout_graph is not None
                                         and self._rollout_graph["key"] == key) else None
        x_in = cached["x_in"] if cached else torch.empty((B, N, U + nc + 2), dtype=torch.float32, device=dev)
        if nc:
            x_in[..., U:U + nc] = x_batch[..., U:U + nc]
        graph, pred_static = (cached["graph"], cached["pred"]) if cached else (None, None)

        def fwd():
            if use_conditional_norm:
                return self.forward(latent_tokens_coord, fixed_coord, x_in[..., :-1], encoder_nbrs=encoder_nbrs,
                                    decoder_nbrs=decoder_nbrs, condition=x_in[..., 0, -2:-1])
            return self.forward(latent_tokens_coord, fixed_coord, x_in, encoder_nbrs=encoder_nbrs, decoder_nbrs=decoder_nbrs)

        for i in range(len(time_indices) - 1):
            t0 = float(t_values[time_indices[i]])
            dt = float(t_values[time_indices[i + 1]]) - t0
            t0n = (t0 - float(stats["start_time"]["mean"])) / float(stats["start_time"]["std"])
            dtn = (dt - float(stats["time_diffs"]["mean"])) / float(stats["time_diffs"]["std"])
            x_in[..., :U] = cur
            x_in[..., U + nc] = t0n
            x_in[..., U + nc + 1] = dtn
            if use_cuda_graph and dev.type == "cuda":
                # one forward captured on the static input buffer and replayed for every rollout step
                if graph is None:
                    side = torch.cuda.Stream()
                    side.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(side):
                        fwd()
                    torch.cuda.current_stream().wait_stream(side)
                    torch.cuda.synchronize()
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph):
                        pred_static = fwd()
                    self._rollout_graph = {"key": key, "graph": graph, "pred": pred_static, "x_in": x_in}
                graph.replay()
                pred = pred_static
            else:
                pred = fwd()
            phys, cur = ops.rollout_update(pred, cur, dt, mode, um, us, sm, ss)
            outs.append(phys)
        return torch.stack(outs, dim=1)
