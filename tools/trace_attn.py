"""Experiment driver: per-phase clock64 trace of attn_bwd_tc_kernel (EXP_TRACE build, tools/build_variants.sh)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from gaot_b200 import ops, _lib
_lib.LIB_PATH = os.path.join(ROOT, "build_exp/lib_EXP_TRACE.so")
B, T, H, hd = 32, 1024, 8, 32
qkv = torch.randn(B * T, 3 * H * hd, device="cuda").to(torch.bfloat16).requires_grad_(True)
go = torch.randn(B * T, H * hd, device="cuda").to(torch.bfloat16)
for _ in range(3):
    o = ops.attention(qkv, B, T, H, H, hd); o.backward(go)
torch.cuda.synchronize()
L = ctypes.CDLL(_lib.LIB_PATH)
buf = (ctypes.c_longlong * (64 * 128))()
print("rc", L.gaot_debug_trace(buf))
for z in (0, 5, 31):
    c = [buf[z * 128 + k] for k in range(64)]            # compute warp (first compute warp of the CTA), 6 slots per step
    m = [buf[z * 128 + 64 + k] for k in range(64)]       # MMA warp, 4 slots per step
    t0 = c[0]
    print(f"CTA z={z}: start->alloc {c[1]-t0}, end {c[62]-t0}, after dealloc {c[63]-t0}")
    for i in range(9):
        a = [c[2 + 6 * i + k] - t0 for k in range(6)]
        print(f"  step {i}: top {a[0]} st_full+{a[1]-a[0]} ld+free+{a[2]-a[1]} done-wait+{a[3]-a[2]} emit+{a[4]-a[3]} drain+{a[5]-a[4]}")
    print("  final", c[2 + 48] - t0)
    for i in range(14):
        a = [m[2 + 4 * i + k] - t0 for k in range(4)]
        print(f"  mma {i}: top {a[0]} st_free-wait+{a[1]-a[0]} issueST+{a[2]-a[1]} pds/dq wait+{a[3]-a[2]}")
