"""Experiment driver: clock64 trace of gemm_bf16_tc_kernel (EXP_TRACE build of tc_gemm.cu, tools/build_variants.sh)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from gaot_b200 import ops, _lib
_lib.LIB_PATH = os.path.join(ROOT, "build_exp/lib_EXP_TRACE.so")
M, N, K, o16 = [int(x) for x in sys.argv[1:5]] if len(sys.argv) > 4 else (32768, 2048, 256, 1)
a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
b = torch.randn(N, K, device="cuda").to(torch.bfloat16)
for _ in range(3):
    ops.gemm_bf16(a, b, M, N, K, out_bf16=bool(o16))
torch.cuda.synchronize()
L = ctypes.CDLL(_lib.LIB_PATH)
buf = (ctypes.c_longlong * (4 * 256))()
print("rc", L.gaot_debug_trace_gemm(buf))
mma = [buf[k] for k in range(256)]; e0 = [buf[256 + k] for k in range(256)]; e1 = [buf[512 + k] for k in range(256)]; tma = [buf[768 + k] for k in range(256)]
t0 = mma[0]
print("peer TMA k-block issue times:", [tma[1 + i] - tma[0] for i in range(24)])
for lt in range(8):
    m = [mma[1 + lt * 8 + k] - t0 for k in range(6)]
    a = [e0[1 + lt * 8 + k] - t0 for k in range(8)]
    b = [e1[1 + lt * 8 + k] - e1[0] for k in range(6)]
    print(f"tile {lt}: MMA top {m[0]} acc_empty+{m[1]-m[0]} kblk {[x - m[1] for x in m[2:6]]} | epi0 top {a[0]} acc_full+{a[1]-a[0]} "
          f"ld0+{a[2]-a[1]} st0+{a[3]-a[2]} ld1+{a[4]-a[3]} st1+{a[5]-a[4]} (cvt+sts {a[6]-a[4]}, fence {a[7]-a[6]}, issue {a[5]-a[7]}) | epi1(peer) top {b[0]} acc_full+{b[1]-b[0]} ld0+{b[2]-b[1]} st0+{b[3]-b[2]} ld1+{b[4]-b[3]} st1+{b[5]-b[4]}")
