"""The training step around the hot path, restated from the reference's trainer loop
(src/utils/optimizers.py:252-255: zero_grad -> train_step -> backward -> step;
src/trainer/static_trainer.py:160-202: forward + MSE; src/core/base_trainer.py:71).

B200-native differences: parameters, gradients and both AdamW moments live in four flat fp32
buffers (one fused optimizer kernel); data parallelism is one process per GPU, the flat gradient
is all-reduced in a few contiguous BUCKETS, each launched on a communication stream as soon as
autograd has produced the last gradient of the bucket, so the reductions run under the rest of
the backward pass (inside the captured CUDA graph when the step is captured).  The reference has
no working DP at all."""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from . import ops


class TrainStep:
    def __init__(self, model: torch.nn.Module, lr: float = 1e-3, weight_decay: float = 1e-2, betas=(0.9, 0.999),
                 eps: float = 1e-8, process_group: Optional[dist.ProcessGroup] = None, distributed: Optional[bool] = None,
                 bucket_bytes: int = 4 << 20):
        self.model = model
        self.lr, self.weight_decay, self.betas, self.eps = lr, weight_decay, betas, eps
        self.group = process_group
        self.distributed = dist.is_available() and dist.is_initialized() if distributed is None else distributed
        self.world = dist.get_world_size(process_group) if self.distributed else 1
        self.params = [p for p in model.parameters() if p.requires_grad]
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        # 16-byte aligned segments so every parameter view stays vector-load friendly
        offs, o = [], 0
        for p in self.params:
            offs.append(o)
            o += (p.numel() + 3) // 4 * 4
        self.flat_param = torch.zeros(o, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(o, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(o, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(o, dtype=torch.float32, device=dev)
        self.grad_views = []
        for p, off in zip(self.params, offs):
            view = self.flat_param[off:off + p.numel()].view_as(p)
            view.copy_(p.data)
            p.data = view
            self.grad_views.append(self.flat_grad[off:off + p.numel()].view_as(p))
            p.grad = None
        self.num_params = n
        self.step_count = 0
        self._buckets, self._bucket_of, self._pending, self._fired, self._work = [], {}, [], [], []
        self._comm = None
        if self.distributed and self.world > 1:
            dist.broadcast(self.flat_param, src=0, group=self.group)     # identical replicas
            self._make_buckets(offs, o, bucket_bytes)

    # ---- bucketed, overlapped gradient all-reduce ---------------------------------------------------
    def _make_buckets(self, offs, total, bucket_bytes):
        """Contiguous parameter ranges of about `bucket_bytes` of the flat gradient, cut from the END of the parameter list
        (autograd produces the decoder's gradients first); one post-accumulate hook per parameter counts a bucket down."""
        lo_idx, hi_idx = len(self.params), len(self.params)
        while hi_idx > 0:
            lo_idx = hi_idx - 1
            while lo_idx > 0 and (offs[hi_idx] if hi_idx < len(offs) else total) - offs[lo_idx - 1] <= bucket_bytes // 4:
                lo_idx -= 1
            flat_hi = offs[hi_idx] if hi_idx < len(offs) else total
            self._buckets.append((lo_idx, hi_idx, offs[lo_idx], flat_hi))
            hi_idx = lo_idx
        for b, (lo, hi, _, _) in enumerate(self._buckets):
            for i in range(lo, hi):
                self._bucket_of[id(self.params[i])] = b
                self.params[i].register_post_accumulate_grad_hook(self._on_grad)
        self._pending = [hi - lo for lo, hi, _, _ in self._buckets]
        self._fired = [False] * len(self._buckets)
        if self.flat_grad.is_cuda:
            self._comm = torch.cuda.Stream()

    def _on_grad(self, p):
        b = self._bucket_of[id(p)]
        self._pending[b] -= 1
        if self._pending[b] == 0 and not self._fired[b]:
            self._fire(b)

    def _fire(self, b):
        """Gradients of bucket b -> its slice of the flat buffer (one multi-tensor copy on the stream that produced them),
        then the all-reduce of that slice on the communication stream."""
        lo, hi, flat_lo, flat_hi = self._buckets[b]
        self._fired[b] = True
        srcs, dsts = [], []
        for p, v in zip(self.params[lo:hi], self.grad_views[lo:hi]):
            if p.grad is None:
                v.zero_()
            else:
                srcs.append(p.grad)
                dsts.append(v)
        if srcs:
            torch._foreach_copy_(dsts, srcs)
        chunk = self.flat_grad[flat_lo:flat_hi]
        if self._comm is not None:
            self._comm.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._comm):
                dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group)
        else:
            self._work.append(dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    # ---- the four calls of optimizers.py:252-255 ---------------------------------------------------
    def zero_grad(self):
        """Gradients are ASSIGNED by autograd (p.grad = None before backward) and gathered into the flat buffer
        by one multi-tensor copy, instead of 60+ per-parameter accumulate kernels into pre-zeroed views."""
        for p in self.params:
            p.grad = None
        if self._buckets:
            self._pending = [hi - lo for lo, hi, _, _ in self._buckets]
            self._fired = [False] * len(self._buckets)

    def gather_gradients(self):
        if self._buckets:                      # data parallel: the hooks have copied (and started reducing) every complete
            for b in range(len(self._buckets)):    # bucket; parameters without a gradient leave theirs to be flushed here
                if not self._fired[b]:
                    self._fire(b)
            return
        srcs, dsts = [], []
        for p, v in zip(self.params, self.grad_views):
            if p.grad is None:
                v.zero_()
            else:
                srcs.append(p.grad)
                dsts.append(v)
        if srcs:
            torch._foreach_copy_(dsts, srcs)

    def reduce_gradients(self):
        """Joins the bucket reductions started during backward (see _fire)."""
        if self._comm is not None:
            torch.cuda.current_stream().wait_stream(self._comm)
        for w in self._work:
            w.wait()
        self._work = []

    def optimizer_step(self):
        self.step_count += 1
        ops.adamw_step(self.flat_param, self.flat_grad, self.exp_avg, self.exp_avg_sq, self.step_count, self.lr,
                       self.betas, self.eps, self.weight_decay, grad_scale=1.0 / self.world)

    def _finish(self, pred, y):
        loss = ops.mse_loss(pred, y)
        loss.backward()
        self.gather_gradients()
        self.reduce_gradients()
        self.optimizer_step()
        return loss.detach()

    def fixed_coords_step(self, latent, coord, x, y, encoder_nbrs=None, decoder_nbrs=None):
        """static_trainer.py:160-178 (+ backward + AdamW): x [B,N,Cin], y [B,N,Cout] -> scalar loss tensor."""
        self.zero_grad()
        pred = self.model(latent_tokens_coord=latent, xcoord=coord, pndata=x, encoder_nbrs=encoder_nbrs,
                          decoder_nbrs=decoder_nbrs)
        return self._finish(pred, y)

    def variable_coords_step(self, latent, x, y, coords, encoder_graphs, decoder_graphs):
        """static_trainer.py:180-202: per-sample coords [B,N,d] and per-sample graphs."""
        self.zero_grad()
        pred = self.model(latent_tokens_coord=latent, xcoord=coords, pndata=x, query_coord=coords,
                          encoder_nbrs=encoder_graphs, decoder_nbrs=decoder_graphs)
        return self._finish(pred, y)

    # ---- CUDA-graph replay of the whole forward + loss + backward ---------------------------------
    def capture_variable_coords(self, latent, x_static, y_static, coords, encoder_graphs, decoder_graphs, warmup=2):
        """vx-mode counterpart of capture_fixed_coords: per-sample coordinates [B,N,d] and per-sample graphs are
        static across replays (one batch of meshes), features/targets are fed through the static buffers."""
        return self._capture(lambda: self.model(latent_tokens_coord=latent, xcoord=coords, pndata=x_static, query_coord=coords,
                                                encoder_nbrs=encoder_graphs, decoder_nbrs=decoder_graphs), y_static, warmup)

    def capture_fixed_coords(self, latent, coord, x_static, y_static, encoder_nbrs=None, decoder_nbrs=None, warmup=2):
        """Capture zero_grad -> forward -> MSE -> backward for STATIC input buffers into one CUDA graph and
        return a `step()` callable that replays it, then runs the fused AdamW kernel (outside the graph: its bias
        correction scalars change per step).  Under data parallelism the bucketed NCCL all-reduces are captured too, on
        their communication stream, overlapping the rest of the backward pass.  New data is fed by
        copying into `x_static` / `y_static` before calling `step()`.  Several hundred small launches per
        step otherwise cost more host time than the GPU needs to execute them."""
        return self._capture(lambda: self.model(latent_tokens_coord=latent, xcoord=coord, pndata=x_static,
                                                encoder_nbrs=encoder_nbrs, decoder_nbrs=decoder_nbrs), y_static, warmup)

    def _capture(self, forward, y_static, warmup):
        from . import _lib

        def fwd_bwd():
            self.zero_grad()
            pred = forward()
            loss = ops.mse_loss(pred, y_static)
            loss.backward()
            self.gather_gradients()
            self.reduce_gradients()                    # joins the communication stream (captured with the graph)
            return loss.detach()

        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):                    # builds graphs/plans/caches, sets kernel attributes
                fwd_bwd()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        with torch.cuda.graph(graph):
            loss_static = fwd_bwd()
        launches = _lib.launch_count() - n0           # kernels of ours inside one replay

        def step():
            graph.replay()
            _lib.note_launches(launches)
            self.optimizer_step()
            return loss_static

        step.graph, step.launches = graph, launches
        return step

    def train_step(self, batch, latent, coord=None):
        """Dispatch on the reference's batch tuple layouts."""
        if len(batch) == 2:
            x, y = batch
            return self.fixed_coords_step(latent, coord, x, y)
        if len(batch) == 5:
            x, y, coords, enc, dec = batch
            return self.variable_coords_step(latent, x, y, coords, enc, dec)
        raise ValueError(f"unsupported batch tuple of length {len(batch)}")


class InputPrefetcher:
    """Feeds the static input buffers of a captured step from pinned host memory without stalling the compute stream:
    `prefetch(host_tensors)` copies the NEXT batch host -> device staging buffers on a dedicated copy stream (it runs under
    the kernels of the current step); `commit()` moves staging -> the graph's static buffers on the compute stream
    (device-to-device, microseconds) right before the next replay.  The reference moves every batch with a blocking
    `.to(device)` inside the step (static_trainer.py:169-170)."""

    def __init__(self, static_tensors):
        self.static = list(static_tensors)
        self.stage = [torch.empty_like(t) for t in self.static]
        self.stream = torch.cuda.Stream()
        self.ready = torch.cuda.Event()
        self.consumed = torch.cuda.Event()
        self.consumed.record(torch.cuda.current_stream())

    def prefetch(self, host_tensors):
        self.stream.wait_event(self.consumed)               # the previous staged batch has been moved out
        with torch.cuda.stream(self.stream):
            for s, h in zip(self.stage, host_tensors):
                s.copy_(h, non_blocking=True)
            self.ready.record(self.stream)

    def commit(self):
        cur = torch.cuda.current_stream()
        cur.wait_event(self.ready)
        for t, s in zip(self.static, self.stage):
            t.copy_(s, non_blocking=True)
        self.consumed.record(cur)
