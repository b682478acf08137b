"""Micro-driver for ncu: tensor-core attention forward + backward at the cfg-2 shape."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from gaot_b200 import ops, _lib  # noqa: E402

if os.environ.get("GAOT_LIB"):          # experiment builds (tools/build_variants.sh)
    _lib.LIB_PATH = os.environ["GAOT_LIB"]

B, T, H, hd = 32, 1024, 8, 32
qkv = torch.randn(B * T, 3 * H * hd, device="cuda").to(torch.bfloat16).requires_grad_(True)
go = torch.randn(B * T, H * hd, device="cuda").to(torch.bfloat16)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    o = ops.attention(qkv, B, T, H, H, hd)
    o.backward(go)
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
e[0].record(); o = ops.attention(qkv, B, T, H, H, hd); e[1].record(); o.backward(go); e[2].record()
torch.cuda.synchronize()
print("fwd us", e[0].elapsed_time(e[1]) * 1e3, "bwd us", e[1].elapsed_time(e[2]) * 1e3)
