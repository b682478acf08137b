"""Times the tensor-core attention forward and backward at the cfg-2 shape (CUDA events around 20 back-to-back calls each;
the backward figure includes its delta / zero-fill / dQ-store helper kernels).  GAOT_B200_ATTN_BWD_QW=128 selects the
one-CTA-per-SM backward kernel."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from gaot_b200 import ops, _lib  # noqa: E402

if os.environ.get("GAOT_LIB"):          # experiment builds (tools/build_variants.sh)
    _lib.LIB_PATH = os.environ["GAOT_LIB"]

B, T, H, hd = 32, 1024, 8, int(sys.argv[1]) if len(sys.argv) > 1 else 32
N = 20
qkv = torch.randn(B * T, 3 * H * hd, device="cuda").to(torch.bfloat16).requires_grad_(True)
go = torch.randn(B * T, H * hd, device="cuda").to(torch.bfloat16)
for _ in range(3):
    o = ops.attention(qkv, B, T, H, H, hd)
    torch.autograd.grad(o, qkv, go)
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
e[0].record()
for _ in range(N):
    o = ops.attention(qkv, B, T, H, H, hd)
e[1].record()
for _ in range(N):
    torch.autograd.grad(o, qkv, go, retain_graph=True)
e[2].record()
torch.cuda.synchronize()
fl = 4.0 * B * H * T * T * hd
tf, tb = e[0].elapsed_time(e[1]) / N * 1e-3, e[1].elapsed_time(e[2]) / N * 1e-3
print(f"hd {hd}: fwd {tf * 1e6:.1f} us ({fl / tf / 1e12:.0f} TFLOP/s)  bwd {tb * 1e6:.1f} us ({2.5 * fl / tb / 1e12:.0f} TFLOP/s)")
