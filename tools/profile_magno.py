"""Micro-driver for ncu: the MAGNO aggregate kernels at BASELINE cfg-2 size (B=32, N=9216, M=4096, C=64).
Usage (GPU box): ncu --set full -k regex:agno ... python tools/profile_magno.py [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from gaot_b200 import ops, synth  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = "cuda"
B, N, C = 32, 9216, 64
x = synth.uniform_points(N, 2, 0).to(dev)
lat = synth.latent_grid((64, 64)).to(dev)
enc = ops.radius_csr(x, lat, 0.033)
dec = ops.radius_csr(lat, x, 0.033)
pe, pd = ops.graph_plan(enc, N), ops.graph_plan(dec, 4096)
f = synth.normal((B, N, C), 1).to(dev).requires_grad_(True)
fl = synth.normal((B, 4096, C), 2).to(dev).requires_grad_(True)
we = synth.normal((pe.E, C), 3).to(dev).requires_grad_(True)
wd = synth.normal((pd.E, C), 4).to(dev).requires_grad_(True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for r in range(reps):
    flush.zero_()
    a = ops.agno_aggregate(f, we, pe)
    flush.zero_()
    b = ops.agno_aggregate(fl, wd, pd)
    flush.zero_()
    (a.sum() * 1.0).backward()
    flush.zero_()
    (b.sum() * 1.0).backward()
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
ts = []
for r in range(10):
    flush.zero_()
    ev[0].record(); a = ops.agno_aggregate(f, we, pe); ev[1].record()
    flush.zero_()
    ev[2].record(); b = ops.agno_aggregate(fl, wd, pd); ev[3].record()
    torch.cuda.synchronize()
    ts.append((ev[0].elapsed_time(ev[1]) * 1e3, ev[2].elapsed_time(ev[3]) * 1e3))
# fused operators (default encoder / decoder paths)
pn = synth.normal((B, N, 1), 5).to(dev)
Wl, bl = synth.normal((C, 1), 6).to(dev), synth.normal((C,), 7).to(dev)
u = synth.normal((1, pd.E, C), 8).to(dev)
tf = []
for r in range(10):
    flush.zero_()
    ev[0].record(); a2 = ops.lift_aggregate(pn, Wl, bl, we.detach(), pe); ev[1].record()
    flush.zero_()
    ev[2].record(); b2 = ops.agno_dot(fl.detach(), u, pd); ev[3].record()
    torch.cuda.synchronize()
    tf.append((ev[0].elapsed_time(ev[1]) * 1e3, ev[2].elapsed_time(ev[3]) * 1e3))
E = pe.E
alg = 4 * C * B * (N + 4096) + 4 * C * E + 8 * E + 8 * 4097
print("E", E, "enc fwd us (L2 flushed) min/med", min(t[0] for t in ts), sorted(t[0] for t in ts)[5],
      "GB/s at min", alg / min(t[0] for t in ts) / 1e3)
print("dec fwd us min/med", min(t[1] for t in ts), sorted(t[1] for t in ts)[5], "GB/s at min", alg / min(t[1] for t in ts) / 1e3)
print("fused lift_aggregate fwd us min/med", min(t[0] for t in tf), sorted(t[0] for t in tf)[5], "alg GB/s at min", alg / min(t[0] for t in tf) / 1e3)
print("fused agno_dot fwd us min/med", min(t[1] for t in tf), sorted(t[1] for t in tf)[5], "alg GB/s at min", alg / min(t[1] for t in tf) / 1e3)
