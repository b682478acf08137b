"""Micro-driver for ncu: the fused MAGNO operators (encoder lift+aggregate, decoder aggregate+read-out; both traversal
orders of the decoder) at BASELINE cfg-2 size (B=32, N=9216, M=4096, C=64), L2 flushed before every launch.
Usage (GPU box): ncu --set full -k regex:"lift_aggregate|agno_dot" -s 14 -c 7 -o gpurun_out/x python tools/profile_magno_fused.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from gaot_b200 import _lib, ops, synth  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
N = int(sys.argv[2]) if len(sys.argv) > 2 else 9216
dev = "cuda"
B, C, M = 32, 64, 4096
x = synth.uniform_points(N, 2, 0).to(dev)
lat = synth.latent_grid((64, 64)).to(dev)
pe, pd = ops.graph_plan(ops.radius_csr(x, lat, 0.033), N), ops.graph_plan(ops.radius_csr(lat, x, 0.033), M)
pn = synth.normal((B, N, 1), 5).to(dev)
pn_t = pn.permute(1, 0, 2).contiguous()
Wl, bl = synth.normal((C, 1), 6).to(dev), synth.normal((C,), 7).to(dev)
we, u = synth.normal((pe.E, C), 3).to(dev), synth.normal((1, pd.E, C), 8).to(dev)
fl = synth.normal((B, M, C), 2).to(dev)
g_enc, g_dec = synth.normal((B, M, C), 17).to(dev), synth.normal((B, N, 1), 18).to(dev)
out_e, out_d = torch.empty((B, M, C), device=dev), torch.empty((B, N, 1), device=dev)
dw, dWl, dbl = torch.empty_like(we), torch.empty_like(Wl), torch.empty_like(bl)
df, du = torch.empty_like(fl), torch.empty_like(u)
ts, tr, tp = pd.transposed()
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
P = lambda t: t.data_ptr()  # noqa: E731
calls = {
    "lift_aggregate_fwd": lambda: _lib.call("gaot_lift_aggregate_fwd", P(pn), P(Wl), P(bl), P(we), P(pe.index), P(pe.splits), P(out_e), B, N, M, C, 1, P(pn_t)),
    "lift_aggregate_bwd": lambda: _lib.call("gaot_lift_aggregate_bwd", P(g_enc), P(pn), P(Wl), P(bl), P(we), P(pe.index), P(pe.splits), P(dw),
                                            P(dWl), P(dbl), B, N, M, pe.E, C, 1, P(pn_t)),
    "agno_dot_fwd_t": lambda: _lib.call("gaot_agno_dot_fwd_t", P(fl), P(u), P(ts), P(tr), P(tp), P(out_d), B, M, N, pd.E, C, 1),
    "agno_dot_bwd_t": lambda: _lib.call("gaot_agno_dot_bwd_t", P(g_dec), P(fl), P(u), P(pd.index), P(pd.splits), P(ts), P(tr), P(tp), P(df), P(du), B, M, N, pd.E, C, 1),
    "agno_dot_fwd": lambda: _lib.call("gaot_agno_dot_fwd", P(fl), P(u), P(pd.index), P(pd.splits), P(out_d), B, M, N, pd.E, C, 1),
    "agno_dot_bwd": lambda: _lib.call("gaot_agno_dot_bwd", P(g_dec), P(fl), P(u), P(pd.index), P(pd.splits), P(ts), P(tr), P(tp), P(df), P(du),
                                      B, M, N, pd.E, C, 1),
}
alg_f = 4 * C * B * (N + M) + 4 * C * pe.E + 8 * pe.E + 8 * (M + 1)
mode = sys.argv[3] if len(sys.argv) > 3 else "write"      # how the 126 MB L2 is flushed before each launch
fl32 = flush.view(torch.int32)
for name, fn in calls.items():
    best = []
    for r in range(reps):
        if mode == "write":
            flush.zero_()                                     # leaves the L2 full of DIRTY lines (evictions cost a write-back)
        elif mode == "read":
            fl32.sum()                                        # leaves the L2 full of CLEAN lines
        else:
            pass                                              # warm: inputs may be L2 resident
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        best.append(e0.elapsed_time(e1) * 1e3)
    print(f"{name:22s} us {sorted(best)[len(best) // 2]:8.1f}  (E enc {pe.E} dec {pd.E})")
