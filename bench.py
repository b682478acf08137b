"""bench.py - mesh-points/sec of one GAOT training step (forward + backward + AdamW) on synthetic meshes.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg2_16k|cfg3|cfg4|cfg5]
                    [--precision bf16|fp32|ffma] [--scaling weak|strong --global-batch G]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...      (one rank per GPU, NCCL)

One "step" = zero_grad -> GAOT.forward -> MSE -> backward -> (gradient all-reduce) -> AdamW over one
batch of synthetic input (reference src/utils/optimizers.py:252-255).  Default workload is BASELINE.json
configs[1]: Poisson-Gauss 2-D, N = 9216 points, latent 64x64, batch 32 PER GPU (weak scaling).
Prints ONE JSON line (rank 0).  `value` = points/s with inputs resident in HBM; `e2e` = the same metric
through the public API with pinned-host inputs copied H2D and the loss read back D2H every step.
`roofline` = the hot kernels timed in isolation (attention backward is the dominant one by the committed ncu launch list);
`fp32_parity_mode` = the same step in the 1e-5 parity mode; `strong_scaling_cfg5` / `rollout_cfg4` = BASELINE configs[4] /
configs[3] as extra blocks of the default line.  `--impl reference` times the reference algorithm's CPU port (oracle/, all
host threads) on the same workload, batch and warm-up.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"          # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line

import torch  # noqa: E402

WORKLOADS = {
    # BASELINE.json configs[1] (hyper-parameters: reference config/examples/time_indep/poisson_gauss.json:12-28)
    "cfg2": dict(d=2, n=9216, latent=(64, 64), B=32, cin=1, cout=1, radius=0.033, C=64, hidden=64, mlp_layers=3,
                 patch=2, dh=256, layers=3, heads=8, desc="Poisson-Gauss 2D fx, N=9216, latent 64x64, B=32/GPU"),
    "cfg2_16k": dict(d=2, n=16384, latent=(64, 64), B=32, cin=1, cout=1, radius=0.033, C=64, hidden=64, mlp_layers=3,
                     patch=2, dh=256, layers=3, heads=8, desc="Poisson-Gauss 2D fx, N=16384, latent 64x64, B=32/GPU"),
    # BASELINE.json configs[2]: airfoil-like 2-D vx mode - every sample has its own mesh and its own radius graphs
    "cfg3": dict(d=2, n=16384, latent=(64, 64), B=64, cin=3, cout=1, radius=0.033, C=64, hidden=64, mlp_layers=3,
                 patch=2, dh=256, layers=3, heads=8, vx=True,
                 desc="NACA0012-like 2D vx (per-sample clustered meshes + per-sample graphs), N=16384, B=64/GPU"),
    "cfg4": dict(d=2, n=16384, latent=(64, 64), B=16, cin=4, cout=2, radius=0.033, C=64, hidden=64, mlp_layers=3,
                 patch=2, dh=256, layers=3, heads=8, desc="ns_gauss-like 2D fx training pair step, N=16384, B=16/GPU"),
    "cfg5": dict(d=3, n=262144, latent=(32, 32, 32), B=1, cin=3, cout=1, radius=0.07, C=48, hidden=64, mlp_layers=3,
                 patch=2, dh=256, layers=3, heads=8, desc="synthetic 3D fx, N=262144, latent 32^3, B=1/GPU"),
}


def make_config(w):
    from gaot_b200.model import AttentionConfig, MAGNOConfig, TransformerConfig
    mc = MAGNOConfig(coord_dim=w["d"], radius=w["radius"], hidden_size=w["hidden"], mlp_layers=w["mlp_layers"],
                     lifting_channels=w["C"], neighbor_search_method="native", precompute_edges=bool(w.get("vx")))
    tc = TransformerConfig(patch_size=w["patch"], hidden_size=w["dh"], num_layers=w["layers"],
                           attn_config=AttentionConfig(num_heads=w["heads"], num_kv_heads=w["heads"]))
    return SimpleNamespace(latent_tokens_size=tuple(w["latent"]), args=SimpleNamespace(magno=mc, transformer=tc))


def synth(w, rank, B):
    from gaot_b200 import synth as cases
    lat = cases.latent_grid(w["latent"])
    if w.get("vx"):
        x = torch.stack([cases.clustered_points(w["n"], w["d"], 100 + 1000 * rank + b) for b in range(B)])
    else:
        x = cases.uniform_points(w["n"], w["d"], 0)
    pn = cases.normal((B, w["n"], w["cin"]), 1000 + rank)
    tgt = cases.normal((B, w["n"], w["cout"]), 2000 + rank)
    return lat, x, pn, tgt


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "20"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return None
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], 0.0, set()
        for r in rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
            except Exception:
                continue
            for nm, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.strip().lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return None
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def load_peaks():
    """Roofline denominators: MEASURED_PEAKS.json (driver-written) or the profiling guide's fallback."""
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        j = json.load(open(pk))
        return {"hbm_gbs": j["hbm_gbs"], "bf16_burst": j["bf16_tflops"], "bf16_sustained": j["bf16_tflops_sustained"],
                "sm_max_mhz": j.get("sm_max_mhz", 1965.0), "src": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "sm_max_mhz": 1965.0, "src": "fallback"}


class Dist:
    """torch.distributed plumbing of one rank (one process per GPU, NCCL)."""

    def __init__(self):
        import torch.distributed as dist
        self.dist = dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py (impl ours) needs a CUDA device: the product has no CPU path")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def timed(self, fn, K):
        """K calls of fn between a barrier + synchronize on both sides; CUDA events; max over ranks (ms)."""
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            fn()
        e1.record()
        self.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        return float(ms)

    def close(self):
        """Barrier, then tear the process group down.  CUDA graphs that captured NCCL work (the bucketed gradient
        all-reduces, the row-sharded MAGNO exchanges) must be gone before the communicator is destroyed, so callers drop
        their captured steps first; a watchdog ends the process if the teardown still blocks (the result line is already out)."""
        if self.world > 1:
            import gc
            import threading
            gc.collect()
            torch.cuda.synchronize()
            self.dist.barrier()
            torch.cuda.synchronize()
            sys.stdout.flush()
            threading.Timer(20.0, lambda: os._exit(0)).start()
            self.dist.destroy_process_group()
            os._exit(0)


class Case:
    """One workload on this rank: model, flat-buffer trainer, synthetic batch, captured step."""

    def __init__(self, D, w, B, precision, use_graph=True, fresh_meshes=0):
        from gaot_b200 import ops as O
        from gaot_b200.model import GAOT
        from gaot_b200.trainer import TrainStep
        self.D, self.w, self.B, self.precision = D, w, B, precision
        dev = D.dev
        O.set_precision(precision)
        torch.manual_seed(42)
        self.model = GAOT(w["cin"], w["cout"], make_config(w)).to(dev)
        self.step = TrainStep(self.model, lr=1e-3, weight_decay=1e-2)
        lat, x, pn, tgt = synth(w, D.rank, B)
        self.lat, self.x, self.pn, self.tgt = lat, x, pn, tgt
        self.lat_d, self.x_d = lat.to(dev), x.to(dev)
        self.pn_d, self.tgt_d = pn.to(dev), tgt.to(dev)
        self.pn_s, self.tgt_s = torch.empty_like(self.pn_d), torch.empty_like(self.tgt_d)
        self.vx = bool(w.get("vx"))
        self.enc_graphs = self.dec_graphs = None
        self.rotate = []
        if self.vx:
            # per-sample graphs built on the GPU by the GraphBuilder drop-in (reference: graph_builder.py:28-87: before training)
            from gaot_b200.datasets import GraphBuilder
            gb = GraphBuilder("native")
            self.enc_graphs, self.dec_graphs = gb.build_graphs_for_split(self.x_d, self.lat_d, w["radius"], [1.0])
            if fresh_meshes:
                # real vx training sees a NEW set of meshes every step: `fresh_meshes` more batches (meshes, graphs, features);
                # the step then runs eagerly and rebuilds the block-diagonal merged graph, its transpose and the geometry
                # statistics for every batch (no CUDA-graph replay of a fixed batch)
                self.rotate = [(self.x_d, self.pn_d, self.tgt_d, self.enc_graphs, self.dec_graphs)]
                for k in range(1, fresh_meshes):
                    _, xk, pnk, tgtk = synth(w, D.rank + 100 * k, B)
                    xk = xk.to(dev)
                    ek, dk = gb.build_graphs_for_split(xk, self.lat_d, w["radius"], [1.0])
                    self.rotate.append((xk, pnk.to(dev), tgtk.to(dev), ek, dk))
                use_graph = False
        self._it = 0
        self.graphed = None
        if use_graph:
            for _ in range(2):
                self.eager_step()
            if self.vx:
                self.graphed = self.step.capture_variable_coords(self.lat_d, self.pn_s, self.tgt_s, self.x_d, self.enc_graphs,
                                                                 self.dec_graphs)
            else:
                self.graphed = self.step.capture_fixed_coords(self.lat_d, self.x_d, self.pn_s, self.tgt_s)
            self.pn_s.copy_(self.pn_d)
            self.tgt_s.copy_(self.tgt_d)

    def eager_step(self):
        from gaot_b200 import ops as O
        O.set_precision(self.precision)
        if self.vx and self.rotate:
            x, pn, tgt, enc, dec = self.rotate[self._it % len(self.rotate)]
            self._it += 1
            # forget the merged plans of earlier batches: a training loader never repeats a combination of samples
            self.model.encoder._merged_cache.clear()
            self.model.decoder._merged_cache.clear()
            return self.step.variable_coords_step(self.lat_d, pn, tgt, x, enc, dec)
        if self.vx:
            return self.step.variable_coords_step(self.lat_d, self.pn_d, self.tgt_d, self.x_d, self.enc_graphs, self.dec_graphs)
        return self.step.fixed_coords_step(self.lat_d, self.x_d, self.pn_d, self.tgt_d)

    def dev_step(self):
        return self.graphed() if self.graphed is not None else self.eager_step()

    def edges(self):
        if self.vx:
            return sum(int(g_[0]["neighbors_index"].numel()) for g_ in self.enc_graphs)
        return int(self.model.encoder.neighbor_cache["graph"][1][0]["neighbors_index"].numel())


def rotating(make, copies):
    """`copies` independent argument sets so that back-to-back launches never find their inputs in the 126 MB L2."""
    return [make(i) for i in range(copies)]


def time_launches(fns, reps):
    """Average device time (us) of one launch: `reps` launches, cycling through the argument sets, between ONE pair of CUDA
    events on the launching stream (torch's current stream is the stream _lib.call launches on)."""
    for f in fns:
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fns[i % len(fns)]()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def isolated_tensor_kernels(w, B, dev, peaks, precision):
    """Each hot tensor-core entry point timed ALONE (20 launches between one event pair, inputs rotated through > 2 L2s of
    memory): attention backward / forward and the three GEMM forms of the largest processor layer (SwiGLU up-projection).
    Algorithmic FLOPs: attention 4 / 10 * B*H*T^2*hd, GEMM 2*M*N*K.  Peak: the measured cuBLAS bf16 BURST
    figure (kernels timed in isolation)."""
    from gaot_b200 import ops as O
    from gaot_b200._lib import call, ptr
    T = 1
    for g_ in w["latent"]:
        T *= g_ // w["patch"]
    H, hd, Dh = w["heads"], w["dh"] // w["heads"], w["dh"]
    F = 4 * Dh
    rows = B * T
    Wq = 3 * H * hd
    scale = 1.0 / hd ** 0.5
    peak = peaks["bf16_burst"]
    out = []
    split = precision != "bf16"
    gen = torch.Generator(device=dev).manual_seed(5)

    def rnd(shape, dtype=torch.float32, s=1.0):
        return (torch.randn(shape, generator=gen, device=dev, dtype=torch.float32) * s).to(dtype)

    def entry(name, kernel, flops, us, note=None):
        d = {"entry": name, "kernel": kernel, "bound": "tensor", "us": round(us, 2), "achieved": round(flops / us / 1e6, 1),
             "peak": peak, "unit": "TFLOP/s", "frac": round(flops / us / 1e6 / peak, 4)}
        if note:
            d["note"] = note
        out.append(d)

    copies = max(2, int(300e6 // (rows * Wq * (4 if split else 2))) + 1)
    if not split:
        sets = []
        for i in range(copies):
            qkv = rnd((rows, Wq), torch.bfloat16)
            o = torch.empty((rows, H * hd), dtype=torch.bfloat16, device=dev)
            lse = torch.empty((B, H, T), device=dev)
            sets.append((qkv, o, lse, rnd((rows, H * hd), torch.bfloat16, 1e-3), torch.empty((B, H, T), device=dev),
                         torch.empty((rows, H * hd), device=dev), torch.empty((rows, Wq), dtype=torch.bfloat16, device=dev)))
        fw = [lambda s_=s_: call("gaot_attention_fwd_bf16", ptr(s_[0]), ptr(s_[1]), ptr(s_[2]), B, T, H, H, hd, Wq, scale) for s_ in sets]
        us = time_launches(fw, 20)
        entry("gaot_attention_fwd_bf16", f"attn_fwd_tc_kernel<{hd}>", 4.0 * B * H * T * T * hd, us)
        bw = [lambda s_=s_: call("gaot_attention_bwd_bf16", ptr(s_[0]), ptr(s_[1]), ptr(s_[2]), ptr(s_[3]), ptr(s_[4]), ptr(s_[5]),
                                 ptr(s_[6]), B, T, H, H, hd, Wq, scale) for s_ in sets]
        us = time_launches(bw, 20)
        entry("gaot_attention_bwd_bf16", f"attn_bwd_tc_kernel<{hd}>", 10.0 * B * H * T * T * hd, us,
              "entry point = delta + memset + attn_bwd_tc_kernel + dq store; the tcgen05 kernel is ~92 % of it")
        del sets, fw, bw
        xs = [rnd((rows, Dh), torch.bfloat16) for _ in range(4)]
        w13 = rnd((2 * F, Dh), torch.bfloat16, Dh ** -0.5)
        ys = [torch.empty((rows, 2 * F), dtype=torch.bfloat16, device=dev) for _ in range(4)]
        us = time_launches([lambda i=i: call("gaot_gemm_bf16", ptr(xs[i]), ptr(w13), ptr(ys[i]), rows, 2 * F, Dh, 1, 0, 0, 0, 2 * F, 0, 0, 0)
                            for i in range(4)], 20)
        entry("gaot_gemm_bf16", "gemm_bf16_tc_kernel<256,0,0,2>", 2.0 * rows * 2 * F * Dh, us, f"forward x W^T, {rows}x{2 * F}x{Dh}, bf16 out")
        dxs = [torch.empty((rows, Dh), dtype=torch.bfloat16, device=dev) for _ in range(4)]
        us = time_launches([lambda i=i: call("gaot_gemm_bf16_nn", ptr(ys[i]), ptr(w13), ptr(dxs[i]), rows, Dh, 2 * F, 1) for i in range(4)], 20)
        entry("gaot_gemm_bf16_nn", "gemm_bf16_tc_kernel<256,0,1,2>", 2.0 * rows * 2 * F * Dh, us, f"input gradient dY W, {rows}x{Dh}x{2 * F}")
        dW = torch.empty((2 * F, Dh), device=dev)
        us = time_launches([lambda i=i: call("gaot_gemm_bf16_tn", ptr(ys[i]), ptr(xs[i]), ptr(dW), 2 * F, Dh, rows, 0) for i in range(4)], 20)
        entry("gaot_gemm_bf16_tn", "gemm_bf16_tc_kernel<256,1,1,2>", 2.0 * rows * 2 * F * Dh, us, f"weight gradient dY^T X, {2 * F}x{Dh}x{rows}")
        # the two GEMMs of the gated FFN that carry the SwiGLU in their epilogues (what the model runs for this layer)
        us_ = [torch.empty((rows, F), dtype=torch.bfloat16, device=dev) for _ in range(4)]
        us = time_launches([lambda i=i: call("gaot_gemm_bf16_swiglu", ptr(xs[i]), ptr(w13), ptr(ys[i]), ptr(us_[i]), rows, 2 * F, Dh)
                            for i in range(4)], 20)
        entry("gaot_gemm_bf16_swiglu", "gemm_bf16_tc_kernel<256,0,0,2> + SwiGLU epilogue", 2.0 * rows * 2 * F * Dh, us,
              f"x W13^T with u = silu(a) * b in the epilogue, {rows}x{2 * F}x{Dh}: writes {rows * 3 * F * 2 / 1e6:.0f} MB")
        w2 = rnd((Dh, F), torch.bfloat16, F ** -0.5)
        dabs = [torch.empty((rows, 2 * F), dtype=torch.bfloat16, device=dev) for _ in range(2)]
        us = time_launches([lambda i=i: call("gaot_gemm_bf16_nn_swiglu_bwd", ptr(xs[i]), ptr(w2), ptr(ys[i]), ptr(dabs[i % 2]), rows, F, Dh)
                            for i in range(4)], 20)
        entry("gaot_gemm_bf16_nn_swiglu_bwd", "gemm_bf16_tc_kernel<256,0,1,2> + SwiGLU backward epilogue", 2.0 * rows * F * Dh, us,
              f"dU = dY W2 consumed in the epilogue, {rows}x{F}x{Dh}: reads and writes {rows * 2 * F * 2 / 1e6:.0f} MB each")
    else:
        sets = []
        for i in range(copies):
            qp, qs = O.split_f16(rnd((rows, Wq)))
            do = rnd((rows, H * hd), s=1e-3)
            dp, ds = O.split_f16(do)
            sets.append((qp, qs, torch.empty((rows, H * hd), device=dev), torch.empty((B, H, T), device=dev), do, dp, ds,
                         torch.empty((B, H, T), device=dev), torch.empty((rows, Wq), device=dev)))
        fw = [lambda s_=s_: call("gaot_attention_fwd_split", ptr(s_[0]), ptr(s_[1]), ptr(s_[2]), ptr(s_[3]), B, T, H, H, hd, Wq, scale)
              for s_ in sets]
        us = time_launches(fw, 20)
        entry("gaot_attention_fwd_split", f"attn_fwd_tc_kernel<{hd},split>", 4.0 * B * H * T * T * hd, us,
              "fp32-equivalent FLOPs; the kernel issues 3x the tensor-core work")
        bw = [lambda s_=s_: call("gaot_attention_bwd_split", ptr(s_[0]), ptr(s_[1]), ptr(s_[2]), ptr(s_[3]), ptr(s_[4]), ptr(s_[5]), ptr(s_[6]),
                                 ptr(s_[7]), ptr(s_[8]), B, T, H, H, hd, Wq, scale) for s_ in sets]
        us = time_launches(bw, 20)
        entry("gaot_attention_bwd_split", f"attn_bwd_tc_kernel<{hd},split>", 10.0 * B * H * T * T * hd, us,
              "fp32-equivalent FLOPs; the kernel issues 3x the tensor-core work")
        del sets, fw, bw
        xs = [O.split_f16(rnd((rows, Dh))) for _ in range(3)]
        wp, wsc = O.split_f16(rnd((2 * F, Dh), s=Dh ** -0.5))
        ys = [torch.empty((rows, 2 * F), device=dev) for _ in range(3)]
        us = time_launches([lambda i=i: O.gemm_split(xs[i][0], xs[i][1], wp, wsc, rows, 2 * F, Dh, out=ys[i]) for i in range(3)], 20)
        entry("gaot_gemm_split", "gemm_bf16_tc_kernel<256,0,0,2,split>", 2.0 * rows * 2 * F * Dh, us,
              f"forward x W^T, {rows}x{2 * F}x{Dh}; fp32-equivalent FLOPs (3 tensor-core terms)")
    return out


def magno_kernels(case, peaks):
    """Each MAGNO gather / reduce operator timed on its own: one launch bracketed by CUDA events on the launching stream,
    preceded by a 512 MB memset that flushes the 126 MB L2; median of 7.  `achieved` uses the ALGORITHMIC bytes SURVEY.md
    section 8d assigns to the operator (gather + weight + segment-sum over [B,N,C] features); `min_gbs` = the bytes the
    fused formulation has to move / time."""
    from gaot_b200 import _lib, ops as O, synth as S
    w, B, dev, model = case.w, case.B, case.D.dev, case.model
    C, N, M = w["C"], w["n"], case.lat.shape[0]
    Cin, Cout = w["cin"], w["cout"]
    enc_g = model.encoder.neighbor_cache["graph"][1][0]
    dec_g = model.decoder.neighbor_cache["graph"][1][0]
    E = int(enc_g["neighbors_index"].numel())

    def alg_fwd(Bc, Ns, Nd):
        return 4.0 * C * Bc * (Ns + Nd) + 4.0 * C * E + 8.0 * E + 8.0 * (Nd + 1)

    def alg_bwd(Bc, Ns, Nd):
        return 4.0 * C * Bc * (2 * Ns + Nd) + 8.0 * C * E + 2 * (8.0 * E + 8.0 * (Nd + 1))

    pe, pd = O.graph_plan(enc_g, N), O.graph_plan(dec_g, M)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    def med_us(fn):
        ts = []
        for _ in range(7):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        return sorted(ts)[3]

    rows = []

    def add(kernel, Ns, Nd, us, alg, min_bytes):
        rows.append(dict(kernel=kernel, bound="hbm", Ns=int(Ns), Nd=int(Nd), us=round(us, 1), achieved=round(alg / us / 1e3, 1),
                         peak=peaks["hbm_gbs"], unit="GB/s", frac=round(alg / us / 1e3 / peaks["hbm_gbs"], 4),
                         alg_bytes=int(alg), min_gbs=round(min_bytes / us / 1e3, 1)))

    Bc = B
    pn_m = S.normal((Bc, N, Cin), 11).to(dev)
    Wl_m, bl_m = S.normal((C, Cin), 12).to(dev), S.normal((C,), 13).to(dev)
    we_m, wd_m = S.normal((pe.E, C), 14).to(dev), S.normal((Cout, pd.E, C), 15).to(dev)
    lat_f = S.normal((Bc, M, C), 16).to(dev)
    g_enc, g_dec = S.normal((Bc, M, C), 17).to(dev), S.normal((Bc, N, Cout), 18).to(dev)
    pn_t_m = pn_m.permute(1, 0, 2).contiguous() if Bc % 8 == 0 else None     # sample-minor copy, as ops.lift_aggregate makes it
    with torch.no_grad():
        t = med_us(lambda: O.lift_aggregate(pn_m, Wl_m, bl_m, we_m, pe))
        add("lift_aggregate_fwd (encoder)", N, M, t, alg_fwd(Bc, N, M), 4.0 * Bc * N * Cin + 4.0 * C * E + 8.0 * E + 4.0 * C * Bc * M)
        ts_, tr_, tp_ = pd.transposed()
        out_d = torch.empty((Bc, N, Cout), device=dev)
        df_, du_ = torch.empty_like(lat_f), torch.empty_like(wd_m)
        dec_min_f = 4.0 * C * Bc * M + 4.0 * C * E * Cout + 8.0 * E + 4.0 * Bc * N * Cout
        dec_min_b = 2 * 4.0 * C * Bc * M + 8.0 * C * E * Cout + 16.0 * E + 4.0 * Bc * N * Cout
        t = med_us(lambda: _lib.call("gaot_agno_dot_fwd_t", lat_f.data_ptr(), wd_m.data_ptr(), ts_.data_ptr(), tr_.data_ptr(),
                                      tp_.data_ptr(), out_d.data_ptr(), Bc, int(M), N, pd.E, C, Cout))
        add("agno_dot_fwd_t (decoder, source walk: default)", M, N, t, alg_fwd(Bc, M, N), dec_min_f)
        dw_, dWl_, dbl_ = torch.empty_like(we_m), torch.empty_like(Wl_m), torch.empty_like(bl_m)
        t = med_us(lambda: _lib.call("gaot_lift_aggregate_bwd", g_enc.data_ptr(), pn_m.data_ptr(), Wl_m.data_ptr(), bl_m.data_ptr(),
                                      we_m.data_ptr(), pe.index.data_ptr(), pe.splits.data_ptr(), dw_.data_ptr(), dWl_.data_ptr(),
                                      dbl_.data_ptr(), Bc, N, int(M), pe.E, C, Cin, pn_t_m.data_ptr() if pn_t_m is not None else 0))
        add("lift_aggregate_bwd (encoder)", N, M, t, alg_bwd(Bc, N, M), 4.0 * Bc * N * Cin + 8.0 * C * E + 8.0 * E + 4.0 * C * Bc * M)
        t = med_us(lambda: _lib.call("gaot_agno_dot_bwd_t", g_dec.data_ptr(), lat_f.data_ptr(), wd_m.data_ptr(), pd.index.data_ptr(),
                                      pd.splits.data_ptr(), ts_.data_ptr(), tr_.data_ptr(), tp_.data_ptr(), df_.data_ptr(), du_.data_ptr(),
                                      Bc, int(M), N, pd.E, C, Cout))
        add("agno_dot_bwd_t (decoder, source walk: default)", M, N, t, alg_bwd(Bc, M, N), dec_min_b)
        # the row-walk forms (deterministic summation order; GAOT_B200_ROW_WALK=1), for comparison
        t = med_us(lambda: _lib.call("gaot_agno_dot_fwd", lat_f.data_ptr(), wd_m.data_ptr(), pd.index.data_ptr(), pd.splits.data_ptr(),
                                      out_d.data_ptr(), Bc, int(M), N, pd.E, C, Cout))
        add("agno_dot_fwd (decoder, row walk)", M, N, t, alg_fwd(Bc, M, N), dec_min_f)
        t = med_us(lambda: _lib.call("gaot_agno_dot_bwd", g_dec.data_ptr(), lat_f.data_ptr(), wd_m.data_ptr(), pd.index.data_ptr(),
                                      pd.splits.data_ptr(), ts_.data_ptr(), tr_.data_ptr(), tp_.data_ptr(), df_.data_ptr(),
                                      du_.data_ptr(), Bc, int(M), N, pd.E, C, Cout))
        add("agno_dot_bwd (decoder, row walk, 2 kernels)", M, N, t, alg_bwd(Bc, M, N), dec_min_b)
        f_m = S.normal((Bc, N, C), 19).to(dev)
        t = med_us(lambda: O.agno_aggregate(f_m, we_m, pe))
        add("agno_aggregate (unfused encoder fwd)", N, M, t, alg_fwd(Bc, N, M), alg_fwd(Bc, N, M))
    del flush
    return rows


def kernel_breakdown(case, dump=""):
    """One extra eager step with CUDA events around every ABI call (same stream).  Host launch gaps are inside these numbers:
    they rank the entry points and count the calls; the roofline figures come from the isolated timings instead."""
    from gaot_b200 import _lib
    rec = []
    _lib.set_profile(rec)
    case.eager_step()
    torch.cuda.synchronize()
    _lib.set_profile(None)
    agg = {}
    for name, a, e0, e1 in rec:
        d = agg.setdefault(name, dict(calls=0, s=0.0))
        d["calls"] += 1
        d["s"] += e0.elapsed_time(e1) * 1e-3
    total = sum(d["s"] for d in agg.values())
    if dump and case.D.rank == 0:
        with open(dump, "w") as fh:
            for name, a, e0, e1 in rec:
                dims = [int(v) for v in a if isinstance(v, int) and 0 < v < 10 ** 9]
                fh.write(f"{e0.elapsed_time(e1) * 1e3:9.1f} us  {name:28s} {dims[:8]}\n")
    return {k: {"calls": d["calls"], "ms": round(d["s"] * 1e3, 3), "share": round(d["s"] / total, 4)}
            for k, d in sorted(agg.items(), key=lambda kv: -kv[1]["s"])}


def strong_scaling_leg(D, args):
    """BASELINE.json configs[4]: synthetic 3-D, N = 262 144, GLOBAL batch 8 sharded over the ranks (strong scaling): every
    rank takes 8 / world samples; coordinates, graphs and the batch-independent edge work are replicated."""
    G = 8
    if G % D.world != 0:
        return {"skipped": f"global batch {G} does not divide over {D.world} ranks"}
    from gaot_b200 import ops as O
    w = WORKLOADS["cfg5"]
    B = G // D.world
    K = 5
    res = {}
    for shard in ((False, True) if D.world > 1 else (False,)):
        # sharded: every rank evaluates 1 / world of the (batch-independent) edge-weight MLP, all-gather / reduce-scatter over
        # NVLink (ops.enable_edge_sharding); replicated: plain data parallelism
        O.enable_edge_sharding(True if shard else None)
        case = Case(D, w, B, args.precision)
        for _ in range(3):
            case.dev_step()
        res[shard] = D.timed(case.dev_step, K) / K
        E = case.edges()
        del case
        torch.cuda.empty_cache()
    O.enable_edge_sharding(None)
    ms = min(res.values())
    out = {"workload": "cfg5", "desc": w["desc"].replace("B=1/GPU", f"global batch {G}"), "scaling": "strong", "global_batch": G,
           "batch_per_gpu": B, "n_gpus": D.world, "ms_per_step": ms, "value": G * w["n"] / (ms * 1e-3), "unit": "points/s",
           "edges": E, "steps": K, "ms_per_step_replicated_edge_mlp": res[False],
           "ms_per_step_sharded_edge_mlp": res.get(True)}
    return out


def rollout_leg(D, args):
    """BASELINE.json configs[3]: ns_gauss-like 2-D time-dependent model, 10-step no-grad autoregressive rollout, batch 16 per
    GPU (GAOT.autoregressive_predict, gaot.py:307-476), one captured forward replayed per step."""
    from gaot_b200 import ops as O
    from gaot_b200.model import GAOT
    w = WORKLOADS["cfg4"]
    B, U = w["B"], w["cout"]
    O.set_precision(args.precision)
    torch.manual_seed(42)
    model = GAOT(w["cin"], w["cout"], make_config(w)).to(D.dev).eval()
    lat, x, pn, _ = synth(w, D.rank, B)
    lat_d, x_d, u0 = lat.to(D.dev), x.to(D.dev), pn[..., :U].contiguous().to(D.dev)
    one = torch.ones(U)
    stats = {"u": {"mean": 0 * one, "std": one}, "der": {"mean": 0 * one, "std": one}, "start_time": {"mean": 0.0, "std": 1.0},
             "time_diffs": {"mean": 0.0, "std": 1.0}}
    tv = torch.linspace(0, 1, 21)[:11]
    idx = list(range(11))

    def run():
        return model.autoregressive_predict(u0, idx, tv, stats, "time_der", latent_tokens_coord=lat_d, fixed_coord=x_d,
                                            use_cuda_graph=True)
    run()
    K = 3
    ms = D.timed(run, K)
    out = {"workload": "cfg4", "desc": "ns_gauss-like 2D fx, 10-step autoregressive rollout (no grad), N=16384, B=16/GPU",
           "rollout_steps": 10, "batch_per_gpu": B, "n_gpus": D.world, "ms_per_rollout": ms / K,
           "value": D.world * B * w["n"] * 10 * K / (ms * 1e-3), "unit": "points/s (mesh points x rollout steps)", "scaling": "weak"}
    del model
    torch.cuda.empty_cache()
    return out


def run_ours(args, w):
    from gaot_b200 import _lib, ops as _ops
    D = Dist()
    world, rank, dev = D.world, D.rank, D.dev
    strong = args.scaling == "strong"
    if strong:
        G = args.global_batch or 8
        if G % world != 0:
            raise SystemExit(f"--global-batch {G} must divide over {world} ranks")
        B = G // world
    else:
        B = args.batch or w["B"]
    peaks = load_peaks()
    if strong and world > 1 and not args.no_row_shard:
        _ops.enable_edge_sharding(True)            # MAGNO rows sharded over the ranks (see gaot_b200/ops.py)
    case = Case(D, w, B, args.precision, use_graph=not args.no_graph, fresh_meshes=args.fresh_meshes if w.get("vx") else 0)
    step, graphed = case.step, case.graphed
    pn_h, tgt_h = case.pn.pin_memory(), case.tgt.pin_memory()
    x_h = case.x.pin_memory() if case.vx else None
    x_stage = torch.empty_like(case.x_d) if case.vx else None

    feeder = None
    if graphed is not None:
        from gaot_b200.trainer import InputPrefetcher
        feeder = InputPrefetcher([case.pn_s, case.tgt_s])
    staged = {"ok": False}

    def e2e_step():
        if feeder is not None:
            # every step's batch goes pinned host -> device inside the timed region; the copy of batch i+1 is issued on the
            # copy stream right after step i's graph has been launched, so it runs under step i's kernels
            if not staged["ok"]:
                feeder.prefetch([pn_h, tgt_h])
            feeder.commit()
            loss = graphed()
            feeder.prefetch([pn_h, tgt_h])
            staged["ok"] = True
            return float(loss)                                                 # .item(): D2H read of the loss
        case.pn_s.copy_(pn_h, non_blocking=True)
        case.tgt_s.copy_(tgt_h, non_blocking=True)
        if case.vx and case.rotate:
            # fresh meshes every step: the coordinates go host -> device as well (graphs are prebuilt, as in the reference)
            x, _, _, enc, dec = case.rotate[case._it % len(case.rotate)]
            case._it += 1
            x_stage.copy_(x_h, non_blocking=True)
            case.model.encoder._merged_cache.clear()
            case.model.decoder._merged_cache.clear()
            loss = step.variable_coords_step(case.lat_d, case.pn_s, case.tgt_s, x, enc, dec)
        elif case.vx:
            loss = step.variable_coords_step(case.lat_d, case.pn_s, case.tgt_s, case.x_d, case.enc_graphs, case.dec_graphs)
        else:
            loss = step.fixed_coords_step(case.lat_d, case.x_d, case.pn_s, case.tgt_s)
        return float(loss)                                                     # .item(): D2H read of the loss

    W = max(args.warmup, 3)
    if case.vx and case.rotate:
        W = max(W, 4 * len(case.rotate))       # every distinct mesh batch four times: the caching allocator keeps growing
                                               # (cudaMalloc inside the step) until it has seen every edge-count pattern
    for _ in range(W):
        case.dev_step()
    sampler = ClockSampler(D.local) if rank == 0 else None
    n0 = _lib.launch_count()
    ms = D.timed(case.dev_step, args.steps)
    launches = _lib.launch_count() - n0
    clocks = sampler.stop() if sampler else None
    e2e_step()
    torch.cuda.synchronize()
    staged["ok"] = False                   # the first timed step copies its own batch inside the timed region
    ms_e2e = D.timed(e2e_step, args.steps)

    # working set of one step vs the L2: activations saved for the backward dominate; measured, not assumed
    torch.cuda.reset_peak_memory_stats()
    m0 = torch.cuda.memory_allocated()
    case.eager_step()
    torch.cuda.synchronize()
    step_bytes = torch.cuda.max_memory_allocated() - m0
    l2_bytes = torch.cuda.get_device_properties(dev).L2_cache_size
    breakdown = kernel_breakdown(case, args.dump_calls)

    E, M = case.edges(), int(case.lat.shape[0])
    total_pts = (B * world) * w["n"] * args.steps
    out = {
        "metric": "mesh-points/sec fwd+bwd", "value": total_pts / (ms * 1e-3), "unit": "points/s", "n_gpus": world,
        "steps": args.steps, "warmup": W, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong" if strong else "weak",
        "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": args.workload, "precision": args.precision, "desc": w["desc"], "batch_per_gpu": B,
                   "global_batch": B * world, "points": w["n"], "latent_tokens": M, "edges": E, "params": step.num_params,
                   "parallelism": f"dp{world}" + ("+magno-rows-sharded" if _ops.edge_shard_world()[0] > 1 else ""),
                   "cuda_graph": graphed is not None,
                   "vx_fresh_mesh_batches": len(case.rotate) if case.vx else None,
                   "l2": f"no flush between steps: one step allocates {step_bytes / 1e6:.0f} MB of activations / workspaces "
                         f"(measured), {step_bytes / max(l2_bytes, 1):.0f}x the {l2_bytes / 1e6:.0f} MB L2"},
        "e2e": {"value": total_pts / (ms_e2e * 1e-3), "unit": "points/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": int(case.pn.numel() * 4 + case.tgt.numel() * 4 + (case.x.numel() * 4 if case.vx and case.rotate else 0)),
                "d2h_bytes_per_step": 4},
        "gpu_launches": int(launches), "clocks": clocks, "kernel_breakdown": breakdown,
    }

    # ---- roofline: the hot kernels timed in isolation on rank 0's GPU (every rank runs them so ranks stay in step) ----
    if not args.no_roofline:
        try:
            tk = isolated_tensor_kernels(w, B, dev, peaks, args.precision)
            mk = magno_kernels(case, peaks) if not case.vx else []
            # dominant kernel by a fixed rule: the attention backward kernel, the largest single kernel of the step in the
            # committed ncu launch list (profiles/README.md section 1); the others are listed beside it
            dom = next(k for k in tk if "attention_bwd" in k["entry"])
            roof = {k: dom[k] for k in ("kernel", "bound", "achieved", "peak", "unit", "frac")}
            roof.update({"entry": dom["entry"], "us": dom["us"], "peak_src": peaks["src"] + " (bf16 burst: kernel timed alone)",
                         "traffic": None, "timing": "20 launches between one CUDA-event pair, inputs rotated through > 2x L2",
                         "share_of_step": round(dom["us"] * 1e-3 * w["layers"] / (ms / args.steps), 4)})
            sms = torch.cuda.get_device_properties(dev).multi_processor_count
            hd = w["dh"] // w["heads"]
            ex2 = dom["achieved"] * 1e12 * dom["us"] * 1e-6 / (10.0 * hd)
            roof["mufu_floor_us"] = round(ex2 / (16.0 * sms * peaks["sm_max_mhz"] * 1e6) * 1e6, 1)
            tr = os.path.join(ROOT, "profiles", "ncu_traffic.json")
            if os.path.exists(tr):
                tj = json.load(open(tr))
                for k in tk + mk:
                    ent = tj.get(k.get("entry") or k["kernel"].split(" ")[0])
                    if ent and args.workload == "cfg2":
                        k["traffic"] = ent["bytes"]
                        k["traffic_src"] = "profiles/" + ent["capture"]
                if dom.get("traffic"):
                    roof["traffic"], roof["traffic_src"] = dom["traffic"], dom["traffic_src"]
            roof["tensor_kernels"] = tk
            roof["magno_hbm"] = mk            # the metric BASELINE.json names: MAGNO gather / scatter GB/s vs the measured HBM peak
            out["roofline"] = roof
            if mk:
                out["roofline_magno"] = {"bound": "hbm", "unit": "GB/s", "peak": peaks["hbm_gbs"], "peak_src": peaks["src"],
                                         "launches": mk}
        except Exception as e:                                   # never lose the headline line to a diagnostics failure
            out["roofline"] = {"error": repr(e)}

    # ---- the same step in the 1e-5 parity mode (split-precision tensor cores), captured the same way ----
    if args.precision == "bf16" and not args.no_fp32_leg:
        try:
            c32 = Case(D, w, B, "fp32", use_graph=not args.no_graph)
            for _ in range(3):
                c32.dev_step()
            ms32 = D.timed(c32.dev_step, 10)
            out["fp32_parity_mode"] = {"value": world * B * w["n"] * 10 / (ms32 * 1e-3), "unit": "points/s", "ms_per_step": ms32 / 10,
                                       "precision": "fp32 (fp16 hi/lo split operands, 3 tcgen05 terms per product, fp32 accumulate)",
                                       "parity": "forward <= 1e-5 rel-L2 vs the reference (tests/test_gpu_model.py)", "cuda_graph": c32.graphed is not None}
            del c32
            torch.cuda.empty_cache()
            _ops.set_precision(args.precision)
        except Exception as e:
            out["fp32_parity_mode"] = {"error": repr(e)}

    # ---- the other BASELINE configurations, as extra blocks of the default line ----
    if args.workload == "cfg2" and not strong and not args.no_extra:
        del case, graphed, feeder
        torch.cuda.empty_cache()
        try:
            out["strong_scaling_cfg5"] = strong_scaling_leg(D, args)
        except Exception as e:
            out["strong_scaling_cfg5"] = {"error": repr(e)}
        try:
            out["rollout_cfg4"] = rollout_leg(D, args)
        except Exception as e:
            out["rollout_cfg4"] = {"error": repr(e)}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = (cpu_reference_unmodified if reference_available(w) else cpu_reference)(w, steps=2, warmup=1, sample_B=B)
    if rank == 0:
        print(json.dumps(out), flush=True)
    case = graphed = feeder = step = None      # captured steps hold NCCL work: release them before the communicator goes
    D.close()


def cpu_reference(w, steps, warmup, sample_B):
    """The reference algorithm's CPU port (oracle/gaot_oracle.py, torch-CPU, all host threads; pinned to the unmodified
    reference by tests/golden) on the workload's mesh and model - the reported baseline, not the target.
    `sample_B` samples per step (the full per-GPU batch unless stated otherwise in `sample`)."""
    from oracle import gaot_oracle as orc
    from gaot_b200.model import GAOT
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(42)
    model = GAOT(w["cin"], w["cout"], make_config(w))       # parameter container only (CPU): same init as our arm
    p = {k: v.detach().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    cfg = orc.make_config(w["latent"], dict(coord_dim=w["d"], radius=w["radius"], hidden_size=w["hidden"], mlp_layers=w["mlp_layers"],
                                            lifting_channels=w["C"]),
                          dict(patch_size=w["patch"], hidden_size=w["dh"], num_layers=w["layers"]),
                          dict(num_heads=w["heads"], num_kv_heads=w["heads"]))
    lat, x, pn, tgt = synth(w, 0, sample_B)
    enc, dec = orc.build_graphs(cfg.args.magno, x, lat), orc.build_graphs(cfg.args.magno, lat, x)
    m = {k: torch.zeros_like(v) for k, v in p.items()}
    v_ = {k: torch.zeros_like(v) for k, v in p.items()}
    times = []
    for it in range(1, warmup + steps + 1):
        t0 = time.perf_counter()
        for t in p.values():
            t.grad = None
        loss = orc.train_step_loss(p, cfg, lat, x, pn, tgt, enc, dec)
        loss.backward()
        with torch.no_grad():
            newp, m, v_ = orc.adamw_step({k: t.detach() for k, t in p.items()}, {k: t.grad for k, t in p.items()}, m, v_, it, 1e-3)
        p = {k: t.requires_grad_(True) for k, t in newp.items()}
        times.append(time.perf_counter() - t0)
    dt = sum(times[warmup:]) / steps
    full = sample_B == w["B"]
    return {"value": sample_B * w["n"] / dt, "unit": "points/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{steps} steps of batch {sample_B} ({'the full per-GPU batch' if full else 'of the workload batch ' + str(w['B'])}) "
                      f"after {warmup} warm-up; graphs prebuilt; {dt:.3f} s/step",
            "s_per_step": dt, "batch": sample_B}


REFERENCE_ROOT = "/root/reference"


def reference_available(w):
    """The unmodified reference can only be timed where its tree exists (the build container; the GPU box has none) and for
    fx workloads (its vx path needs the graph containers of its own data pipeline)."""
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "model")) and not w.get("vx")


def cpu_reference_unmodified(w, steps, warmup, sample_B):
    """The UNMODIFIED reference model (`src.model.gaot.GAOT` imported from /root/reference, CPU, fp32, all host threads) in the
    reference's own step - zero_grad, forward, MSE, backward, torch.optim.AdamW (optimizers.py:252-255, base_trainer.py:71) -
    on the workload's mesh and batch.  torch_scatter / omegaconf / rotary_embedding_torch are absent from the image; the
    import stand-ins of oracle/ref_shims (documented there) replace them, hence kind "reference+shims"."""
    for pth in (os.path.join(ROOT, "oracle", "ref_shims"), REFERENCE_ROOT):
        if pth not in sys.path:
            sys.path.insert(0, pth)
    from src.model.gaot import GAOT as RefGAOT                      # the reference
    from src.model.layers.attn import AttentionConfig as RAC, TransformerConfig as RTC
    from src.model.layers.magno import MAGNOConfig as RMC
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(42)
    cfg = SimpleNamespace(latent_tokens_size=tuple(w["latent"]), args=SimpleNamespace(
        magno=RMC(coord_dim=w["d"], radius=w["radius"], hidden_size=w["hidden"], mlp_layers=w["mlp_layers"], lifting_channels=w["C"],
                  neighbor_search_method="native"),
        transformer=RTC(patch_size=w["patch"], hidden_size=w["dh"], num_layers=w["layers"],
                        attn_config=RAC(num_heads=w["heads"], num_kv_heads=w["heads"]))))
    model = RefGAOT(w["cin"], w["cout"], cfg)
    model.train()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=1e-2)
    lat, x, pn, tgt = synth(w, 0, sample_B)
    times = []
    for _ in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad()
        loss = torch.nn.functional.mse_loss(model(lat, x, pn), tgt)
        loss.backward()
        opt.step()
        times.append(time.perf_counter() - t0)
    dt = sum(times[warmup:]) / steps
    return {"value": sample_B * w["n"] / dt, "unit": "points/s", "cores": torch.get_num_threads(), "kind": "reference+shims",
            "sample": f"{steps} steps of batch {sample_B} of the unmodified reference (src.model.gaot.GAOT + torch.optim.AdamW) after "
                      f"{warmup} warm-up (the first builds and caches the graphs); {dt:.3f} s/step",
            "s_per_step": dt, "batch": sample_B}


def run_reference(args, w):
    """Reference arm: the reference's own CPU implementation of the path (the oracle port: the reference is pure Python
    with no installable package), all host threads, on OUR arm's config - the same workload, the same
    per-GPU batch, the same warm-up count.  The step count is bounded so the run ends within a few minutes."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    B = args.batch or w["B"]
    W = max(args.warmup, 3)
    steps = args.steps
    # the unmodified reference where its tree exists (the build container), else its pinned port (the GPU box has no /root/reference)
    ref_fn = cpu_reference_unmodified if reference_available(w) else cpu_reference
    base = ref_fn(w, steps=1, warmup=1, sample_B=B)                 # probe: seconds per step on this box
    budget = 150.0                                                  # seconds for the timed region + warm-up
    fit = max(1, int(budget / max(base["s_per_step"], 1e-3)) - W)
    timed_steps = min(steps, fit)
    base = ref_fn(w, steps=timed_steps, warmup=W, sample_B=B)
    if timed_steps < steps:
        base["sample"] += f"; {timed_steps} of the requested {steps} steps timed (bounded to ~{budget:.0f} s), per-step time reported"
    out = {"impl": "reference", "metric": "mesh-points/sec fwd+bwd", "value": base["value"], "unit": "points/s",
           "n_gpus": world, "steps": steps, "warmup": W, "ms_per_step": base["s_per_step"] * 1e3,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": args.workload, "precision": "fp32", "desc": w["desc"], "batch_per_gpu": B, "global_batch": B,
                      "points": w["n"], "parallelism": "cpu (rank 0 only; the reference has no working data parallelism)"},
           "cpu_baseline": base,
           "e2e": {"value": base["value"], "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch override (weak scaling)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="strong: --global-batch samples are sharded over the ranks (BASELINE configs[4]: cfg5, global batch 8)")
    ap.add_argument("--global-batch", type=int, default=0)
    ap.add_argument("--no-row-shard", action="store_true",
                    help="strong scaling: plain data parallelism (edge-weight MLP replicated) instead of row-sharded MAGNO")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fp32-leg", action="store_true", help="skip the extra fp32-parity-mode timing")
    ap.add_argument("--no-roofline", action="store_true", help="skip the isolated kernel timings")
    ap.add_argument("--no-extra", action="store_true", help="skip the cfg5 strong-scaling and cfg4 rollout blocks of the default line")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying a CUDA graph")
    ap.add_argument("--fresh-meshes", type=int, default=3,
                    help="vx workloads: number of distinct mesh batches cycled through, a new one every step (eager step, merged "
                         "graphs rebuilt per batch); 0 = replay one fixed batch through a CUDA graph")
    ap.add_argument("--dump-calls", default="", help="write the per-ABI-call timing list of one step to this file")
    ap.add_argument("--precision", default=os.environ.get("GAOT_B200_PRECISION", "bf16"), choices=["fp32", "ffma", "bf16"],
                    help="processor numerics: bf16 tensor cores, fp32 = split-precision tensor cores (parity), ffma = fp32 FFMA")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w)
    else:
        run_ours(args, w)


if __name__ == "__main__":
    main()
