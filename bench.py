"""bench.py - mesh-points/sec of one GAOT training step (forward + backward + AdamW) on synthetic meshes.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg2_16k|cfg4|cfg5]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...      (one rank per GPU, NCCL)

One "step" = zero_grad -> GAOT.forward -> MSE -> backward -> (gradient all-reduce) -> AdamW over one
batch of synthetic input (reference src/utils/optimizers.py:252-255).  Default workload is BASELINE.json
configs[1]: Poisson-Gauss 2-D, N = 9216 points, latent 64x64, batch 32 PER GPU (weak scaling).
Prints ONE JSON line (rank 0).  `value` = points/s with inputs resident in HBM; `e2e` = the same metric
through the public API with pinned-host inputs copied H2D and the loss read back D2H every step.
`--impl reference` times the reference algorithm's CPU port (oracle/, all host threads) on a bounded
sample of the same workload.
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


# algorithmic work per ABI call (DESIGN.md section "kernels"): (flops, bytes) from the call's own arguments
def call_work(name, a):
    if name == "gaot_gemm_f32":
        M, N, K = a[3], a[4], a[5]
        return 2.0 * M * N * K, 4.0 * (M * K + N * K + M * N)
    if name == "gaot_gemm_bf16":
        M, N, K = a[3], a[4], a[5]
        return 2.0 * M * N * K, 2.0 * (M * K + N * K) + (2.0 if a[6] else 4.0) * M * N
    if name == "gaot_gemm_split":
        M, N, K = a[3], a[4], a[5]
        return 2.0 * M * N * K, 4.0 * (M * K + N * K + M * N)
    if name == "gaot_attention_fwd_split":
        B, T, H, hd = a[4], a[5], a[6], a[8]
        return 4.0 * B * H * T * T * hd, 4.0 * B * T * H * hd * 4
    if name == "gaot_attention_bwd_split":
        B, T, H, hd = a[9], a[10], a[11], a[13]
        return 10.0 * B * H * T * T * hd, 4.0 * B * T * H * hd * 8
    if name in ("gaot_gemm_bf16_nn", "gaot_gemm_bf16_tn"):
        M, N, K = a[3], a[4], a[5]
        return 2.0 * M * N * K, 2.0 * (M * K + N * K) + 4.0 * M * N
    if name == "gaot_attention_fwd":
        B, T, H, hd = a[5], a[6], a[7], a[9]
        return 4.0 * B * H * T * T * hd, 4.0 * B * T * H * hd * 4
    if name == "gaot_attention_fwd_bf16":
        B, T, H, hd = a[3], a[4], a[5], a[7]
        return 4.0 * B * H * T * T * hd, 2.0 * B * T * H * hd * 4
    if name == "gaot_attention_bwd_bf16":
        B, T, H, hd = a[7], a[8], a[9], a[11]
        return 10.0 * B * H * T * T * hd, 2.0 * B * T * H * hd * 8
    if name == "gaot_attention_bwd":
        B, T, H, hd = a[10], a[11], a[12], a[14]
        return 10.0 * B * H * T * T * hd, 4.0 * B * T * H * hd * 8
    return None, None


def run_ours(args, w):
    import torch.distributed as dist
    from gaot_b200 import _lib
    from gaot_b200.model import GAOT
    from gaot_b200.trainer import TrainStep

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl ours) needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch or w["B"]
    from gaot_b200 import ops as _ops
    _ops.set_precision(args.precision)
    torch.manual_seed(42)
    model = GAOT(w["cin"], w["cout"], make_config(w)).to(dev)
    step = TrainStep(model, lr=1e-3, weight_decay=1e-2)
    lat, x, pn, tgt = synth(w, rank, B)
    lat_d, x_d = lat.to(dev), x.to(dev)
    pn_d, tgt_d = pn.to(dev), tgt.to(dev)
    pn_h, tgt_h = pn.pin_memory(), tgt.pin_memory()
    pn_s, tgt_s = torch.empty_like(pn_d), torch.empty_like(tgt_d)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, K):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    vx = bool(w.get("vx"))
    enc_graphs = dec_graphs = None
    if vx:      # per-sample graphs built on the GPU in one batched launch each (reference: graph_builder.py:28-87)
        from gaot_b200 import ops as _o
        enc_graphs = [[g_] for g_ in _o.radius_csr(x_d, lat_d, w["radius"])]
        dec_graphs = [[g_] for g_ in _o.radius_csr(lat_d, x_d, w["radius"])]

    def eager_step():
        if vx:
            return step.variable_coords_step(lat_d, pn_d, tgt_d, x_d, enc_graphs, dec_graphs)
        return step.fixed_coords_step(lat_d, x_d, pn_d, tgt_d)

    if args.no_graph:
        graphed = None
    else:
        for _ in range(2):
            eager_step()
        if vx:
            graphed = step.capture_variable_coords(lat_d, pn_s, tgt_s, x_d, enc_graphs, dec_graphs)
        else:
            graphed = step.capture_fixed_coords(lat_d, x_d, pn_s, tgt_s)       # static input buffers pn_s / tgt_s
        pn_s.copy_(pn_d); tgt_s.copy_(tgt_d)

    def dev_step():
        return graphed() if graphed is not None else eager_step()

    feeder = None
    if graphed is not None:
        from gaot_b200.trainer import InputPrefetcher
        feeder = InputPrefetcher([pn_s, tgt_s])
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
        pn_s.copy_(pn_h, non_blocking=True)
        tgt_s.copy_(tgt_h, non_blocking=True)
        if graphed is not None:
            loss = graphed()
        elif vx:
            loss = step.variable_coords_step(lat_d, pn_s, tgt_s, x_d, enc_graphs, dec_graphs)
        else:
            loss = step.fixed_coords_step(lat_d, x_d, pn_s, tgt_s)
        return float(loss)                                                     # .item(): D2H read of the loss

    W = max(args.warmup, 3)
    for _ in range(W):
        dev_step()
    sampler = ClockSampler(local) if rank == 0 else None
    n0 = _lib.launch_count()
    ms = timed(dev_step, args.steps)
    launches = _lib.launch_count() - n0
    clocks = sampler.stop() if sampler else None
    e2e_step()
    torch.cuda.synchronize()
    staged["ok"] = False                   # the first timed step copies its own batch inside the timed region
    ms_e2e = timed(e2e_step, args.steps)

    # per-kernel breakdown: one extra step with CUDA events around every ABI call (same stream)
    rec = []
    _lib.set_profile(rec)
    eager_step()
    torch.cuda.synchronize()
    _lib.set_profile(None)
    agg = {}
    for name, a, e0, e1 in rec:
        t = e0.elapsed_time(e1) * 1e-3
        fl, by = call_work(name, a)
        d = agg.setdefault(name, dict(calls=0, s=0.0, flops=0.0, bytes=0.0, items=[]))
        d["calls"] += 1; d["s"] += t
        if fl:
            d["flops"] += fl
        d["items"].append((t, a))
    total_s = sum(d["s"] for d in agg.values())
    if args.dump_calls and rank == 0:
        with open(args.dump_calls, "w") as fh:
            for name, a, e0, e1 in rec:
                dims = [int(v) for v in a if isinstance(v, int) and 0 < v < (1 << 40) and v < 10 ** 9]
                fh.write(f"{e0.elapsed_time(e1) * 1e3:9.1f} us  {name:28s} {dims[:8]}\n")

    peaks = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "src": "fallback"}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        j = json.load(open(pk))
        peaks = {"hbm_gbs": j["hbm_gbs"], "bf16_tflops": j["bf16_tflops_sustained"], "src": "measured"}

    # MAGNO scatter (the metric BASELINE.json names).  `achieved` uses the ALGORITHMIC bytes SURVEY.md section 8d
    # assigns to the operator (gather + weight + segment-sum over [B,N,C] features); the fused kernels do that
    # job while touching fewer bytes (`min_bytes`: what the fused formulation has to move), see DESIGN.md section 4.
    if vx:
        E = sum(int(g_[0]["neighbors_index"].numel()) for g_ in enc_graphs)
    else:
        enc_g = model.encoder.neighbor_cache["graph"][1][0]
        E = int(enc_g["neighbors_index"].numel())
    C, N, M = w["C"], w["n"], lat.shape[0]
    Cin, Cout = w["cin"], w["cout"]

    def alg_fwd(Bc, Ns, Nd):
        return 4.0 * C * Bc * (Ns + Nd) + 4.0 * C * E + 8.0 * E + 8.0 * (Nd + 1)

    def alg_bwd(Bc, Ns, Nd):
        return 4.0 * C * Bc * (2 * Ns + Nd) + 8.0 * C * E + 2 * (8.0 * E + 8.0 * (Nd + 1))

    # Each MAGNO gather/reduce operator is timed on its own: one launch bracketed by CUDA events on the launching
    # stream, preceded by a 512 MB memset that flushes the 126 MB L2 and keeps the GPU busy while the host enqueues
    # (the per-call numbers of the eager breakdown step include host launch gaps that are comparable to these
    # 30-60 us kernels).  Median of 7.
    from gaot_b200 import ops as O, synth as S
    magno = []
    if not vx:
        dec_g = model.decoder.neighbor_cache["graph"][1][0]
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

        Bc = B
        pn_m = S.normal((Bc, N, Cin), 11).to(dev)
        Wl_m, bl_m = S.normal((C, Cin), 12).to(dev), S.normal((C,), 13).to(dev)
        we_m, wd_m = S.normal((pe.E, C), 14).to(dev), S.normal((Cout, pd.E, C), 15).to(dev)
        lat_f = S.normal((Bc, M, C), 16).to(dev)
        g_enc, g_dec = S.normal((Bc, M, C), 17).to(dev), S.normal((Bc, N, Cout), 18).to(dev)
        with torch.no_grad():
            t = med_us(lambda: O.lift_aggregate(pn_m, Wl_m, bl_m, we_m, pe))
            magno.append(dict(kernel="lift_aggregate_fwd (encoder)", Ns=N, Nd=int(M), us=t, gbs=alg_fwd(Bc, N, M) / t / 1e3,
                              min_bytes=4.0 * Bc * N * Cin + 4.0 * C * E + 8.0 * E + 4.0 * C * Bc * M))
            t = med_us(lambda: O.agno_dot(lat_f, wd_m, pd))
            magno.append(dict(kernel="agno_dot_fwd (decoder)", Ns=int(M), Nd=N, us=t, gbs=alg_fwd(Bc, M, N) / t / 1e3,
                              min_bytes=4.0 * C * Bc * M + 4.0 * C * E * Cout + 8.0 * E + 4.0 * Bc * N * Cout))
            dw_, dWl_, dbl_ = torch.empty_like(we_m), torch.empty_like(Wl_m), torch.empty_like(bl_m)
            t = med_us(lambda: _lib.call("gaot_lift_aggregate_bwd", g_enc.data_ptr(), pn_m.data_ptr(), Wl_m.data_ptr(), bl_m.data_ptr(),
                                          we_m.data_ptr(), pe.index.data_ptr(), pe.splits.data_ptr(), dw_.data_ptr(), dWl_.data_ptr(),
                                          dbl_.data_ptr(), Bc, N, int(M), pe.E, C, Cin))
            magno.append(dict(kernel="lift_aggregate_bwd (encoder)", Ns=N, Nd=int(M), us=t, gbs=alg_bwd(Bc, N, M) / t / 1e3,
                              min_bytes=4.0 * Bc * N * Cin + 8.0 * C * E + 8.0 * E + 4.0 * C * Bc * M))
            ts_, tr_, tp_ = pd.transposed()
            df_, du_ = torch.empty_like(lat_f), torch.empty_like(wd_m)
            t = med_us(lambda: _lib.call("gaot_agno_dot_bwd", g_dec.data_ptr(), lat_f.data_ptr(), wd_m.data_ptr(), pd.index.data_ptr(),
                                          pd.splits.data_ptr(), ts_.data_ptr(), tr_.data_ptr(), tp_.data_ptr(), df_.data_ptr(),
                                          du_.data_ptr(), Bc, int(M), N, pd.E, C, Cout))
            magno.append(dict(kernel="agno_dot_bwd (decoder, 2 kernels)", Ns=int(M), Nd=N, us=t, gbs=alg_bwd(Bc, M, N) / t / 1e3,
                              min_bytes=2 * 4.0 * C * Bc * M + 8.0 * C * E * Cout + 16.0 * E + 4.0 * Bc * N * Cout))
            f_m = S.normal((Bc, N, C), 19).to(dev)
            t = med_us(lambda: O.agno_aggregate(f_m, we_m, pe))
            magno.append(dict(kernel="agno_aggregate (unfused encoder fwd)", Ns=N, Nd=int(M), us=t, gbs=alg_fwd(Bc, N, M) / t / 1e3,
                              min_bytes=alg_fwd(Bc, N, M)))
        del flush
    best = max(magno, key=lambda r: r["us"]) if magno else None
    top_name = max(agg, key=lambda k: agg[k]["s"])
    top = agg[top_name]
    if top["flops"] > 0:
        ach = top["flops"] / top["s"] / 1e12
        roof = {"kernel": top_name, "bound": "tensor", "achieved": round(ach, 3), "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                "frac": round(ach / peaks["bf16_tflops"], 5), "traffic": None, "peak_src": peaks["src"],
                "share_of_step": round(top["s"] / total_s, 3),
                "note": ("tcgen05 kernel" if "bf16" in top_name else "fp32 FFMA parity-path kernel measured against the bf16 tensor peak")}
        if top_name.startswith("gaot_attention") and "bf16" in top_name:
            # at head dim 32 the softmax exponentials, not the MMAs, bound attention: one ex2 per score on the MUFU pipe
            # (16 / clk / SM on sm_100a) against 4*hd (forward) or 10*hd (backward) tensor FLOPs per score
            a0 = top["items"][0][1]
            hd_ = a0[7] if top_name.endswith("fwd_bf16") else a0[11]
            per = (4.0 if top_name.endswith("fwd_bf16") else 10.0) * hd_
            ex2s = top["flops"] / per
            floor_s = ex2s / (16.0 * 148 * 1965e6)
            roof["mufu_bound"] = {"ex2": ex2s, "floor_ms": round(floor_s * 1e3, 4), "measured_ms": round(top["s"] * 1e3, 4),
                                  "frac_of_mufu_floor": round(floor_s / top["s"], 4),
                                  "note": "MUFU.EX2 floor at 16/clk/SM, 148 SMs, 1965 MHz; tensor floor is lower at hd=%d" % hd_}
    else:
        ach = best["gbs"] if best else 0.0
        roof = {"kernel": top_name, "bound": "hbm", "achieved": round(ach, 1), "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": round(ach / peaks["hbm_gbs"], 4), "traffic": None, "peak_src": peaks["src"],
                "share_of_step": round(top["s"] / total_s, 3)}
    tr = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tr):
        ent = json.load(open(tr)).get(top_name)
        if ent and args.workload == "cfg2":
            roof["traffic"] = ent["bytes"]                 # ncu --set full capture of this kernel at this shape
            roof["traffic_src"] = "profiles/" + ent["capture"] if ent["capture"].startswith("raw/") else ent["capture"]
    roof_magno = None
    if magno:
        roof_magno = {"bound": "hbm", "unit": "GB/s", "peak": peaks["hbm_gbs"], "peak_src": peaks["src"],
                      "bytes": "algorithmic bytes of the unfused operator (SURVEY 8d); min_gbs = bytes the fused kernel must move / time",
                      "launches": [dict(kernel=r["kernel"], Ns=r["Ns"], Nd=r["Nd"], us=round(r["us"], 1), achieved=round(r["gbs"], 1),
                                        frac=round(r["gbs"] / peaks["hbm_gbs"], 4), min_gbs=round(r["min_bytes"] / r["us"] / 1e3, 1))
                                   for r in magno],
                      "share_of_step": round(sum(agg[k]["s"] for k in agg if "aggregate" in k or "agno_dot" in k) / total_s, 4),
                      "traffic": None}
    breakdown = {k: {"calls": d["calls"], "ms": round(d["s"] * 1e3, 3), "share": round(d["s"] / total_s, 4)}
                 for k, d in sorted(agg.items(), key=lambda kv: -kv[1]["s"])}
    for k, d in agg.items():                           # tensor-pipe kernels: achieved TFLOP/s of the eager-timed calls
        if d["flops"] > 0:                             # (host launch gaps included, so a lower bound) vs the measured bf16 peak
            breakdown[k]["tflops"] = round(d["flops"] / d["s"] / 1e12, 1)
            breakdown[k]["frac_of_bf16_peak"] = round(d["flops"] / d["s"] / 1e12 / peaks["bf16_tflops"], 4)

    pts = world * B * w["n"] * args.steps
    out = {
        "metric": "mesh-points/sec fwd+bwd", "value": pts / (ms * 1e-3), "unit": "points/s", "n_gpus": world,
        "steps": args.steps, "warmup": W, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": args.workload, "precision": args.precision, "desc": w["desc"], "batch_per_gpu": B, "points": w["n"], "latent_tokens": int(M),
                   "edges": E, "params": step.num_params, "parallelism": f"dp{world}", "cuda_graph": graphed is not None,
                   "l2": "no flush: per-step working set (~3 GB of activations at B=32) exceeds the 126 MB L2"},
        "e2e": {"value": pts / (ms_e2e * 1e-3), "unit": "points/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": int(pn.numel() * 4 + tgt.numel() * 4), "d2h_bytes_per_step": 4},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "roofline_magno": roof_magno,
        "kernel_breakdown": breakdown,
    }
    if args.precision == "bf16" and not args.no_fp32_leg:
        # the same step in the fp32 parity mode (the mode the 1e-5 parity tests run in), eager launches
        _ops.set_precision("fp32")
        for _ in range(3):
            eager_step()
        ms32 = timed(eager_step, 5)
        _ops.set_precision(args.precision)
        out["fp32_parity_mode"] = {"value": world * B * w["n"] * 5 / (ms32 * 1e-3), "unit": "points/s", "ms_per_step": ms32 / 5}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_reference(w, steps=2, warmup=1, sample_B=min(8, w["B"]))
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cpu_reference(w, steps, warmup, sample_B):
    """The reference algorithm's CPU port (oracle/gaot_oracle.py, torch-CPU, all host threads) on a bounded
    sample (sample_B samples of the same mesh/model) - the reported baseline, not the target."""
    from oracle import cases, gaot_oracle as orc
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
    return {"value": sample_B * w["n"] / dt, "unit": "points/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{steps} steps of batch {sample_B} (of the workload's batch) after {warmup} warm-up; graphs prebuilt; "
                      f"{dt:.3f} s/step", "s_per_step": dt}


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    sample_B = min(8, w["B"])
    base = cpu_reference(w, steps=args.steps, warmup=max(1, min(args.warmup, 2)), sample_B=sample_B)
    out = {"impl": "reference", "metric": "mesh-points/sec fwd+bwd", "value": base["value"], "unit": "points/s",
           "n_gpus": world, "steps": args.steps, "warmup": max(1, min(args.warmup, 2)), "ms_per_step": base["s_per_step"] * 1e3,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": args.workload, "desc": w["desc"], "batch_per_gpu": args.batch or w["B"], "points": w["n"],
                      "parallelism": "cpu"},
           "cpu_baseline": base,
           "e2e": {"value": base["value"], "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch override")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fp32-leg", action="store_true", help="skip the extra fp32-parity-mode timing")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying a CUDA graph")
    ap.add_argument("--dump-calls", default="", help="write the per-ABI-call timing list of one step to this file")
    ap.add_argument("--precision", default=os.environ.get("GAOT_B200_PRECISION", "bf16"), choices=["fp32", "ffma", "bf16"],
                    help="processor numerics: fp32 parity path or bf16 tensor-core path")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w)
    else:
        run_ours(args, w)


if __name__ == "__main__":
    main()
