"""Micro-driver for the fused edge-weight kernels (gaot_edge_mlp_{fwd,bwd}): the cfg-3 sized merged graph of 64 clustered
meshes (E ~ 3.6 M edges), event timing and an ncu target:
    ncu --set full -k regex:edge_mlp -s 2 -c 2 -o gpurun_out/x python tools/profile_edge_mlp.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from gaot_b200 import ops, synth  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = "cuda"
lat = synth.latent_grid((64, 64)).to(dev)
xs = torch.stack([synth.clustered_points(16384, 2, 100 + b) for b in range(S)]).to(dev)
graphs = ops.radius_csr(xs, lat, 0.033)
plan = ops.merge_graphs(graphs, 16384)
y, x = xs.reshape(-1, 2), lat.repeat(S, 1)
gen = torch.Generator().manual_seed(0)
dims = [4, 64, 64, 64, 64]
params = []
for a, b in zip(dims[:-1], dims[1:]):
    params += [(torch.randn(b, a, generator=gen) / a ** 0.5).to(dev).requires_grad_(True), (0.1 * torch.randn(b, generator=gen)).to(dev).requires_grad_(True)]
alpha = torch.rand(plan.E, generator=gen).to(dev)
gw = torch.randn(plan.E, 64, generator=gen).to(dev)
for it in range(3):
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    w = ops.edge_mlp(y, x, plan, alpha, params)
    e1.record()
    w.backward(gw)
    e2.record()
    torch.cuda.synchronize()
    print(f"E = {plan.E}: fwd {e0.elapsed_time(e1):.3f} ms  bwd {e1.elapsed_time(e2):.3f} ms  "
          f"({plan.E / e0.elapsed_time(e1) / 1e6:.2f} / {plan.E / e1.elapsed_time(e2) / 1e6:.2f} G edges/s)")
