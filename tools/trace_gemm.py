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
buf = (ctypes.c_longlong * (8 * 256))()
print("rc", L.gaot_debug_trace_gemm(buf))
for cta, base in ((0, 0), (77, 1024)):
    tma = [buf[base + k] for k in range(256)]
    mma = [buf[base + 256 + k] for k in range(256)]
    ep = [buf[base + 512 + k] for k in range(256)]
    ep9 = [buf[base + 768 + k] for k in range(256)]
    t0 = tma[0]
    print(f"CTA {cta}: alloc done +{tma[1]-t0}")
    print("  TMA k-block issue times:", [tma[2 + i] - t0 for i in range(24)])
    for lt in range(6):
        m = [mma[2 + lt * 8 + k] - t0 for k in range(6)]
        e = [ep[2 + lt * 12 + k] - t0 for k in range(11)]
        print(f"  tile {lt}: MMA top {m[0]} acc_empty +{m[1]-m[0]} kblocks ready {[x - m[1] for x in m[2:6]]} | "
              f"epi(w2) top {e[0]} acc_full +{e[1]-e[0]} chunk0: ld +{e[2]-e[1]} sts +{e[3]-e[2]} lds +{e[4]-e[3]} st +{e[5]-e[4]} "
              f"chunk1: ld +{e[6]-e[5]} sts +{e[7]-e[6]} lds +{e[8]-e[7]} st +{e[9]-e[8]} end +{e[10]-e[9]}")
