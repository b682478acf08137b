"""Times the narrow fp32 GEMM kernels (weights-in-registers forms of gaot_gemm_f32) at the geometry-embedding shapes of
cfg-2: CUDA events around 50 back-to-back launches."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from gaot_b200 import ops, _lib  # noqa: E402

if os.environ.get("GAOT_LIB"):
    _lib.LIB_PATH = os.environ["GAOT_LIB"]

dev = "cuda"
for M, N, K in ((9216, 64, 7), (9216, 64, 64), (4096, 64, 64), (9216, 1, 64), (131072, 64, 64)):
    x = torch.randn(M, K, device=dev)
    W = torch.randn(N, K, device=dev)
    g = torch.randn(M, N, device=dev)
    def run(fn):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 50 * 1e3
    t_f = run(lambda: ops.gemm(x, W, M, N, K, K, K, 0, 1))            # y = x W^T        (rows mode, BT)
    t_dx = run(lambda: ops.gemm(g, W, M, K, N, N, K, 0, 0))           # dx = dy W        (rows mode)
    t_dw = run(lambda: ops.gemm(g, x, N, K, M, N, K, 1, 0))           # dW = dy^T x      (reduce mode)
    print(f"M={M} N={N} K={K}: fwd {t_f:.1f} us  dx {t_dx:.1f} us  dW {t_dw:.1f} us")
