"""Times gaot_rmsnorm_fwd / gaot_rmsnorm_bwd at the cfg-2 processor shape ([32768, 256], 'both' mode): 50 launches per event pair,
inputs rotated through > 2x the L2."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from gaot_b200 import _lib  # noqa: E402

if os.environ.get("GAOT_LIB"):
    _lib.LIB_PATH = os.environ["GAOT_LIB"]
from gaot_b200._lib import call, ptr  # noqa: E402

rows, D, dev = 32768, 256, "cuda"
n = 10
xs = [torch.randn(rows, D, device=dev) for _ in range(n)]
dys = [torch.randn(rows, D, device=dev) for _ in range(n)]
d16 = [torch.randn(rows, D, device=dev).to(torch.bfloat16) for _ in range(n)]
dxs = [torch.empty(rows, D, device=dev) for _ in range(n)]
g = torch.ones(D, device=dev)
rstd = torch.rand(rows, device=dev) + 0.5
dg = torch.zeros(D, device=dev)
y = torch.empty(rows, D, device=dev)
y16 = torch.empty(rows, D, device=dev, dtype=torch.bfloat16)


def run(fn):
    for i in range(n):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(50):
        fn(i % n)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 50 * 1e3


print("rmsnorm_bwd (dy fp32 + dy16):", round(run(lambda i: call("gaot_rmsnorm_bwd", ptr(dys[i]), ptr(d16[i]), ptr(xs[i]), ptr(g), ptr(rstd), ptr(dxs[i]), ptr(dg), rows, D)), 1), "us  (117 MB)")
print("rmsnorm_bwd (dy16 only):     ", round(run(lambda i: call("gaot_rmsnorm_bwd", 0, ptr(d16[i]), ptr(xs[i]), ptr(g), ptr(rstd), ptr(dxs[i]), ptr(dg), rows, D)), 1), "us  (84 MB)")
print("rmsnorm_fwd (both outputs):  ", round(run(lambda i: call("gaot_rmsnorm_fwd", ptr(xs[i]), ptr(g), ptr(y), ptr(y16), ptr(rstd), rows, D, 1e-6)), 1), "us  (84 MB)")
