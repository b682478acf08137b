"""Micro-benchmark of the tcgen05 GEMM at the processor's shapes (GPU box)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from gaot_b200 import ops  # noqa: E402

dev = "cuda"
shapes = [(32768, 2048, 256, True), (32768, 768, 256, True), (32768, 256, 1024, False), (32768, 256, 256, False)]
for M, N, K, o16 in shapes:
    a = torch.randn(M, K, device=dev).to(torch.bfloat16)
    b = torch.randn(N, K, device=dev).to(torch.bfloat16)
    for _ in range(3):
        ops.gemm_bf16(a, b, M, N, K, out_bf16=o16)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.gemm_bf16(a, b, M, N, K, out_bf16=o16)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    print(f"M={M} N={N} K={K} out_bf16={o16}: {us:.1f} us  {2.0 * M * N * K / us / 1e6:.0f} TFLOP/s")

# dW form (contraction over the 32768 rows, split-K)
for M, N, R in [(2048, 256, 32768), (768, 256, 32768), (256, 256, 32768), (256, 1024, 32768)]:
    a = torch.randn(R, M, device=dev).to(torch.bfloat16)
    b = torch.randn(R, N, device=dev).to(torch.bfloat16)
    for _ in range(3):
        ops.gemm_bf16_tn(a, b, M, N, R)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.gemm_bf16_tn(a, b, M, N, R)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    print(f"tn M={M} N={N} R={R}: {us:.1f} us  {2.0 * M * N * R / us / 1e6:.0f} TFLOP/s")
