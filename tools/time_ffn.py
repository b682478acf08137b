"""Times the gated FFN of one processor layer at the cfg-2 shape (M = 32768, D = 256, F = 1024), forward + backward, fused
SwiGLU epilogues (ops.swiglu_ffn) against the unfused chain (linear -> swiglu -> linear): CUDA events around 20 iterations."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from gaot_b200 import ops  # noqa: E402

M, D, F = 32768, 256, 1024
dev = "cuda"
ops.set_precision("bf16")
x = torch.randn(M, D, device=dev).to(torch.bfloat16).requires_grad_(True)
w1, w3 = [(torch.randn(F, D, device=dev) / 16).requires_grad_(True) for _ in range(2)]
w2 = (torch.randn(D, F, device=dev) / 32).requires_grad_(True)
res = torch.randn(M, D, device=dev)
gy = torch.randn(M, D, device=dev)


def fused():
    y = ops.swiglu_ffn(x, ops.interleave_w13(w1, w3), w2, add=res)
    y.backward(gy)


def unfused():
    ab = ops.linear(x, torch.cat([w1, w3], 0), tc=True, out_bf16=True)
    y = ops.linear(ops.swiglu(ab), w2, None, add=res, tc=True)
    y.backward(gy)


for name, fn in (("fused", fused), ("unfused", unfused)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per FFN forward + backward")
