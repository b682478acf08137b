"""Experiment driver: per-phase clock64 trace of attn_fwd_tc_kernel (EXP_TRACE_FWD build, tools/build_variants.sh attn_tc EXP_TRACE_FWD)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from gaot_b200 import ops, _lib
_lib.LIB_PATH = os.path.join(ROOT, "build_exp/lib_EXP_TRACE_FWD.so")
B, T, H, hd = 32, 1024, 8, 32
qkv = torch.randn(B * T, 3 * H * hd, device="cuda").to(torch.bfloat16)
for _ in range(3):
    o = ops.attention(qkv, B, T, H, H, hd)
torch.cuda.synchronize()
L = ctypes.CDLL(_lib.LIB_PATH)
buf = (ctypes.c_longlong * (4 * 128))()
print("rc", L.gaot_debug_trace_fwd(buf))
for name, base in (("z=0 warp2", 0), ("z=0 warp6", 128), ("z=20 warp2", 256)):
    c = [buf[base + k] for k in range(128)]
    t0 = c[0]
    print(name)
    for i in range(8):
        a = [c[i * 8 + k] - t0 for k in range(8)]
        print(f"  tile {i}: top {a[0]} s_full+{a[1]-a[0]} ld+{a[2]-a[1]} max/xchg+{a[3]-a[2]} rescale+{a[4]-a[3]} p_free+{a[5]-a[4]} exp+{a[6]-a[5]} fence+{a[7]-a[6]}")
