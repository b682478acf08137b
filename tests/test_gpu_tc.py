"""GPU tests of the tensor-core (tcgen05 / TMA / TMEM) performance path.  Run with -m gpu.

Tolerances: the GEMM itself is compared against an fp64 product of the SAME bf16-rounded operands, so
only accumulation order differs (<= 2e-5 relative L2, K up to 32768).  The bf16 *mode* of the model
rounds activations and weights to bf16 (8 mantissa bits) at every processor GEMM input, so it is
checked against the fp64 reference golden vectors at 2e-2 (forward) / 6e-2 (gradients) relative L2;
the 1e-5 parity bar applies to the fp32 mode (tests/test_gpu_model.py).

The split-precision kernels of that fp32 mode (fp16 hi + lo planes, three tensor-core terms per product) are tested here
kernel by kernel against fp64 of the ORIGINAL fp32 operands at 2e-6 (GEMM) / 5e-6 (attention)."""
import pytest
import torch

from oracle import cases
from tests.helpers import build_model, golden_params, load_golden, rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 64, 256), (1000, 200, 136), (4096, 768, 256), (256, 256, 32768),
                                   (77, 24, 8), (2048, 2048, 1024)])
def test_gemm_bf16_vs_fp64_of_rounded_operands(M, N, K):
    from gaot_b200 import ops
    a = cases.normal((M, K), 1).to(DEV)
    b = (cases.normal((N, K), 2) / K ** 0.5).to(DEV)
    bias = cases.normal((N,), 3).to(DEV)
    R = M // 2 if M % 2 == 0 else M
    add = cases.normal((R, N), 4).to(DEV)
    ab, _ = ops.cast_bf16(a)
    bb, bt = ops.cast_bf16(b, want=True, want_t=True)
    assert torch.equal(ab, a.to(torch.bfloat16)) and torch.equal(bt, b.to(torch.bfloat16).T.contiguous())
    assert torch.equal(ops.transpose_bf16(ab), ab.T.contiguous())
    want = ab.double() @ bb.double().T
    got = ops.gemm_bf16(ab, bb, M, N, K)
    assert rel_l2(got, want) < 2e-5
    got = ops.gemm_bf16(ab, bb, M, N, K, bias=bias, add=add, add_rows=R)
    assert rel_l2(got, want + bias.double() + add.double().repeat(M // R, 1)) < 2e-5
    got16 = ops.gemm_bf16(ab, bb, M, N, K, out_bf16=True)
    assert got16.dtype == torch.bfloat16 and rel_l2(got16, want) < 4e-3
    acc = torch.ones((M, N), device=DEV)
    ops.gemm_bf16(ab, bb, M, N, K, out=acc, accumulate=True)
    assert rel_l2(acc, want + 1.0) < 2e-5


def test_linear_tc_fwd_bwd():
    from gaot_b200 import ops
    M, N, K = 1024, 256, 512
    x, W, b = cases.normal((M, K), 5), cases.normal((N, K), 6) / K ** 0.5, cases.normal((N,), 7)
    add, gy = cases.normal((M, N), 8), cases.normal((M, N), 9)
    ops.set_precision("bf16")
    try:
        xs = [t.clone().to(DEV).requires_grad_(True) for t in (x, W, b, add)]
        y = ops.linear(xs[0], xs[1], xs[2], add=xs[3], tc=True)
        y.backward(gy.to(DEV))
    finally:
        ops.set_precision("fp32")
    xr = [t.clone().double().requires_grad_(True) for t in (x, W, b, add)]
    yr = xr[0] @ xr[1].T + xr[2] + xr[3]
    yr.backward(gy.double())
    assert rel_l2(y, yr) < 5e-3
    for a, r_ in zip(xs, xr):
        assert rel_l2(a.grad, r_.grad) < 6e-3


@pytest.mark.parametrize("name", ["tiny2d", "tiny3d"])
def test_model_bf16_mode_close_to_reference(name):
    from gaot_b200 import ops
    c = cases.CASES[name]
    g = load_golden(name)
    ops.set_precision("bf16")
    try:
        model = build_model(c)
        model.load_state_dict({k: v.to(DEV) for k, v in golden_params(g).items()})
        lat, x, pn, tgt = [t.to(DEV) for t in cases.case_inputs(name)]
        out = model(lat, x, pn)
        loss = ops.mse_loss(out, tgt)
        loss.backward()
    finally:
        ops.set_precision("fp32")
    assert rel_l2(out, g["out"]) < 2e-2
    # all gradients together (tensors with near-zero gradients, e.g. q/k projections of an almost uniform
    # attention, are dominated by rounding noise individually and get a looser per-tensor bound)
    mine = torch.cat([p.grad.flatten().double().cpu() for _, p in model.named_parameters()])
    ref = torch.cat([torch.from_numpy(g["grad." + k]).flatten().double() for k, _ in model.named_parameters()])
    assert rel_l2(mine, ref) < 3e-2
    gmax = max(float(torch.from_numpy(g["grad." + k]).norm()) for k, _ in model.named_parameters())
    for k, p in model.named_parameters():
        if float(torch.from_numpy(g["grad." + k]).norm()) > 1e-3 * gmax:
            assert rel_l2(p.grad, g["grad." + k]) < 0.1, k


@pytest.mark.parametrize("B,T,H,Hkv,hd", [(2, 256, 2, 2, 32), (1, 1024, 8, 8, 32), (2, 200, 4, 2, 32), (1, 384, 2, 2, 64), (1, 130, 2, 1, 16),
                                          (3, 40, 2, 2, 32), (2, 72, 2, 1, 16)])
def test_attention_fwd_bf16_vs_fp64(B, T, H, Hkv, hd):
    from gaot_b200 import _lib
    from gaot_b200._lib import call, ptr
    W = (H + 2 * Hkv) * hd
    qkv = cases.normal((B * T, W), 21).to(DEV).to(torch.bfloat16)
    o = torch.empty((B * T, H * hd), dtype=torch.bfloat16, device=DEV)
    lse = torch.empty((B, H, T), dtype=torch.float32, device=DEV)
    call("gaot_attention_fwd_bf16", ptr(qkv), ptr(o), ptr(lse), B, T, H, Hkv, hd, W, 1.0 / hd ** 0.5)
    torch.cuda.synchronize()
    x = qkv.double().cpu()
    q = x[:, :H * hd].view(B, T, H, hd).transpose(1, 2)
    k = x[:, H * hd:(H + Hkv) * hd].view(B, T, Hkv, hd).transpose(1, 2).repeat_interleave(H // Hkv, dim=1)
    v = x[:, (H + Hkv) * hd:].view(B, T, Hkv, hd).transpose(1, 2).repeat_interleave(H // Hkv, dim=1)
    s = q @ k.transpose(-1, -2) / hd ** 0.5
    want = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B * T, H * hd)
    assert rel_l2(lse, torch.logsumexp(s, -1)) < 2e-6
    assert rel_l2(o, want) < 8e-3            # P and O are rounded to bf16


@pytest.mark.parametrize("B,T,H,Hkv,hd", [(2, 256, 2, 2, 32), (1, 1024, 8, 8, 32), (2, 200, 4, 2, 32), (1, 384, 2, 2, 64), (1, 130, 2, 1, 16),
                                          (3, 40, 2, 2, 32), (2, 72, 2, 1, 16)])
def test_attention_bwd_bf16_vs_fp64(B, T, H, Hkv, hd):
    from gaot_b200 import ops
    W = (H + 2 * Hkv) * hd
    qkv = cases.normal((B * T, W), 22).to(torch.bfloat16)
    go = cases.normal((B * T, H * hd), 23).to(torch.bfloat16)
    a = qkv.to(DEV).requires_grad_(True)
    o = ops.attention(a, B, T, H, Hkv, hd)
    o.backward(go.to(DEV))
    x = qkv.double().requires_grad_(True)
    q = x[:, :H * hd].view(B, T, H, hd).transpose(1, 2)
    k = x[:, H * hd:(H + Hkv) * hd].view(B, T, Hkv, hd).transpose(1, 2).repeat_interleave(H // Hkv, dim=1)
    v = x[:, (H + Hkv) * hd:].view(B, T, Hkv, hd).transpose(1, 2).repeat_interleave(H // Hkv, dim=1)
    want = (torch.softmax(q @ k.transpose(-1, -2) / hd ** 0.5, -1) @ v).transpose(1, 2).reshape(B * T, H * hd)
    want.backward(go.double())
    assert rel_l2(o, want) < 8e-3
    gq, gk, gv = a.grad[:, :H * hd], a.grad[:, H * hd:(H + Hkv) * hd], a.grad[:, (H + Hkv) * hd:]
    rq, rk, rv = x.grad[:, :H * hd], x.grad[:, H * hd:(H + Hkv) * hd], x.grad[:, (H + Hkv) * hd:]
    assert rel_l2(gv, rv) < 1e-2, "dV"
    assert rel_l2(gk, rk) < 2e-2, "dK"
    assert rel_l2(gq, rq) < 2e-2, "dQ"


@pytest.mark.parametrize("M,N,R", [(256, 256, 32768), (2048, 256, 4096), (768, 256, 1000), (64, 64, 512), (200, 72, 136)])
def test_gemm_bf16_tn_vs_fp64(M, N, R):
    """dW form: C = A^T B with the contraction over the rows of both operands (MN-major tcgen05 operands)."""
    from gaot_b200 import ops
    a = cases.normal((R, M), 41).to(DEV).to(torch.bfloat16)
    b = (cases.normal((R, N), 42) / R ** 0.5).to(DEV).to(torch.bfloat16)
    want = a.double().T @ b.double()
    got = ops.gemm_bf16_tn(a, b, M, N, R)
    assert rel_l2(got, want) < 2e-5
    acc = torch.ones((M, N), device=DEV)
    ops.gemm_bf16_tn(a, b, M, N, R, out=acc, accumulate=True)
    assert rel_l2(acc, want + 1.0) < 2e-5


def test_swiglu_bf16():
    from gaot_b200 import ops
    ab = cases.normal((300, 128), 43).to(torch.bfloat16)
    go = cases.normal((300, 64), 44).to(torch.bfloat16)
    a = ab.double().requires_grad_(True)
    want = torch.nn.functional.silu(a[:, :64]) * a[:, 64:]
    want.backward(go.double())
    b = ab.to(DEV).requires_grad_(True)
    got = ops.swiglu(b)
    got.backward(go.to(DEV))
    assert got.dtype == torch.bfloat16 and rel_l2(got, want) < 4e-3 and rel_l2(b.grad, a.grad) < 6e-3


@pytest.mark.parametrize("M,N,K", [(32768, 256, 768), (4096, 1024, 256), (1000, 72, 136), (128, 64, 64)])
def test_gemm_bf16_nn_vs_fp64(M, N, K):
    """dX form: C = A B with B [K,N] row-major consumed as an MN-major tcgen05 operand."""
    from gaot_b200 import ops
    a = cases.normal((M, K), 51).to(DEV).to(torch.bfloat16)
    b = (cases.normal((K, N), 52) / K ** 0.5).to(DEV).to(torch.bfloat16)
    want = a.double() @ b.double()
    assert rel_l2(ops.gemm_bf16_nn(a, b, M, N, K), want) < 2e-5
    assert rel_l2(ops.gemm_bf16_nn(a, b, M, N, K, out_bf16=True), want) < 4e-3


def test_linear_tc_with_activation():
    from gaot_b200 import ops
    M, N, K = 5000, 64, 64
    x, W, b, gy = cases.normal((M, K), 61), cases.normal((N, K), 62) / K ** 0.5, cases.normal((N,), 63), cases.normal((M, N), 64)
    ops.set_precision("bf16")
    try:
        xs = [t.clone().to(DEV).requires_grad_(True) for t in (x, W, b)]
        y = ops.linear(xs[0], xs[1], xs[2], act=ops.ACT_GELU, tc=True, out_bf16=True)
        assert y.dtype == torch.bfloat16
        y.backward(gy.to(DEV).to(torch.bfloat16))
    finally:
        ops.set_precision("fp32")
    xr = [t.clone().double().requires_grad_(True) for t in (x, W, b)]
    yr = torch.nn.functional.gelu(xr[0] @ xr[1].T + xr[2])
    yr.backward(gy.double())
    assert rel_l2(y, yr) < 8e-3
    for a, r_ in zip(xs, xr):
        assert rel_l2(a.grad, r_.grad) < 1.5e-2


# ------------------------------------------------------------------------------------------------
# split-precision (parity) tensor-core kernels
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("scale", [1.0, 3e-7, 4e5])
def test_split_f16_planes_reconstruct_fp32(scale):
    """x * s = hi + lo to ~2^-22 of the tensor maximum, s a power of two with max|x| * s in [2^13, 2^14)."""
    from gaot_b200 import ops
    x = (cases.normal((300, 136), 70) * scale).to(DEV)
    planes, sc = ops.split_f16(x)
    s, inv = float(sc[0]), float(sc[1])
    assert s * inv == 1.0 and abs(torch.log2(torch.tensor(s)).item() - round(torch.log2(torch.tensor(s)).item())) == 0
    amax = float(x.abs().max())
    assert 2 ** 13 <= amax * s < 2 ** 14
    back = (planes[0].double() + planes[1].double()) * inv
    assert float((back - x.double()).abs().max()) <= 2.0 ** -22 * amax
    z = torch.zeros((8, 8), device=DEV)
    pz, scz = ops.split_f16(z)
    assert float(scz[0]) == 1.0 and float(pz.abs().max()) == 0.0


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 64, 256), (1000, 200, 136), (4096, 768, 256), (77, 24, 8),
                                   (2048, 1024, 1024), (32768, 256, 256)])
def test_gemm_split_forward_form_vs_fp64(M, N, K):
    from gaot_b200 import ops
    a = cases.normal((M, K), 71).to(DEV)
    b = (cases.normal((N, K), 72) / K ** 0.5).to(DEV)
    bias = cases.normal((N,), 73).to(DEV)
    R = M // 2 if M % 2 == 0 else M
    add = cases.normal((R, N), 74).to(DEV)
    (ap, sa), (bp, sb) = ops.split_f16(a), ops.split_f16(b)
    want = a.double() @ b.double().T
    assert rel_l2(ops.gemm_split(ap, sa, bp, sb, M, N, K), want) < 2e-6
    got = ops.gemm_split(ap, sa, bp, sb, M, N, K, bias=bias, add=add, add_rows=R)
    assert rel_l2(got, want + bias.double() + add.double().repeat(M // R, 1)) < 2e-6
    pre = torch.empty((M, N), device=DEV)
    got = ops.gemm_split(ap, sa, bp, sb, M, N, K, bias=bias, act=ops.ACT_GELU, pre=pre)
    assert rel_l2(pre, want + bias.double()) < 2e-6
    assert rel_l2(got, torch.nn.functional.gelu(want + bias.double())) < 2e-6
    acc = torch.ones((M, N), device=DEV)
    ops.gemm_split(ap, sa, bp, sb, M, N, K, out=acc, accumulate=True)
    assert rel_l2(acc, want + 1.0) < 2e-6


@pytest.mark.parametrize("M,N,K", [(32768, 256, 768), (4096, 1024, 256), (1000, 72, 136), (128, 64, 64)])
def test_gemm_split_input_gradient_form_vs_fp64(M, N, K):
    """C = A B with B [K, N] planes read MN-major (dX = dY W); A carries gradient-sized values (1e-6)."""
    from gaot_b200 import ops
    a = (cases.normal((M, K), 75) * 1e-6).to(DEV)
    b = (cases.normal((K, N), 76) / K ** 0.5).to(DEV)
    (ap, sa), (bp, sb) = ops.split_f16(a), ops.split_f16(b)
    assert rel_l2(ops.gemm_split(ap, sa, bp, sb, M, N, K, b_mn=True), a.double() @ b.double()) < 2e-6


@pytest.mark.parametrize("M,N,R", [(256, 256, 32768), (2048, 256, 4096), (768, 256, 1000), (64, 64, 512), (200, 72, 136)])
def test_gemm_split_weight_gradient_form_vs_fp64(M, N, R):
    """C = A^T B, contraction over the rows of both operands (dW = dY^T X), split-K with TMA reduce-add."""
    from gaot_b200 import ops
    a = (cases.normal((R, M), 77) * 1e-6).to(DEV)
    b = cases.normal((R, N), 78).to(DEV)
    (ap, sa), (bp, sb) = ops.split_f16(a), ops.split_f16(b)
    assert rel_l2(ops.gemm_split(ap, sa, bp, sb, M, N, R, a_mn=True, b_mn=True), a.double().T @ b.double()) < 2e-6


def test_linear_split_fwd_bwd_vs_fp64():
    from gaot_b200 import ops
    M, N, K = 1024, 256, 512
    x, W, b = cases.normal((M, K), 5), cases.normal((N, K), 6) / K ** 0.5, cases.normal((N,), 7)
    add, gy = cases.normal((M, N), 8), cases.normal((M, N), 9) * 1e-5
    assert ops.get_precision() == "fp32"
    xs = [t.clone().to(DEV).requires_grad_(True) for t in (x, W, b, add)]
    y = ops.linear(xs[0], xs[1], xs[2], add=xs[3], act=ops.ACT_GELU, tc=True)
    y.backward(gy.to(DEV))
    xr = [t.clone().double().requires_grad_(True) for t in (x, W, b, add)]
    yr = torch.nn.functional.gelu(xr[0] @ xr[1].T + xr[2] + xr[3])
    yr.backward(gy.double())
    assert rel_l2(y, yr) < 2e-6
    for a, r_ in zip(xs, xr):
        assert rel_l2(a.grad, r_.grad) < 3e-6


def _attn_fp64(x, go, B, T, H, Hkv, hd):
    q = x[:, :H * hd].view(B, T, H, hd).transpose(1, 2)
    k = x[:, H * hd:(H + Hkv) * hd].view(B, T, Hkv, hd).transpose(1, 2).repeat_interleave(H // Hkv, dim=1)
    v = x[:, (H + Hkv) * hd:].view(B, T, Hkv, hd).transpose(1, 2).repeat_interleave(H // Hkv, dim=1)
    s = q @ k.transpose(-1, -2) / hd ** 0.5
    o = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B * T, H * hd)
    if go is not None:
        o.backward(go)
    return o, torch.logsumexp(s, -1)


@pytest.mark.parametrize("gscale", [1.0, 1e-6])
@pytest.mark.parametrize("B,T,H,Hkv,hd", [(2, 256, 2, 2, 32), (1, 1024, 8, 8, 32), (2, 200, 4, 2, 32), (1, 130, 2, 1, 16),
                                          (3, 64, 2, 2, 16)])
def test_attention_split_fwd_bwd_vs_fp64(B, T, H, Hkv, hd, gscale):
    """Split-precision tcgen05 attention (gaot_attention_{fwd,bwd}_split through ops.attention in 'fp32' mode) against fp64
    of the original fp32 inputs; sharp softmax rows (scores x3) and gradient-sized upstream gradients (1e-6) included."""
    from gaot_b200 import ops
    W = (H + 2 * Hkv) * hd
    qkv = cases.normal((B * T, W), 81)
    qkv[:, :H * hd] *= 3.0                                  # sharper rows: exercises the lazy rescaling
    go = cases.normal((B * T, H * hd), 82) * gscale
    assert ops.get_precision() == "fp32"
    a = qkv.to(DEV).requires_grad_(True)
    o = ops.attention(a, B, T, H, Hkv, hd)
    assert o.dtype == torch.float32
    o.backward(go.to(DEV))
    x = qkv.double().requires_grad_(True)
    want, _ = _attn_fp64(x, go.double(), B, T, H, Hkv, hd)
    assert rel_l2(o, want) < 5e-6
    gq, gk, gv = a.grad[:, :H * hd], a.grad[:, H * hd:(H + Hkv) * hd], a.grad[:, (H + Hkv) * hd:]
    rq, rk, rv = x.grad[:, :H * hd], x.grad[:, H * hd:(H + Hkv) * hd], x.grad[:, (H + Hkv) * hd:]
    assert rel_l2(gv, rv) < 5e-6, "dV"
    assert rel_l2(gk, rk) < 1e-5, "dK"
    assert rel_l2(gq, rq) < 1e-5, "dQ"


# ------------------------------------------------------------------------------------------------
# bf16 mode at model level on every fixture family (reference goldens), with stated tolerances
# ------------------------------------------------------------------------------------------------
BF16_FWD_TOL, BF16_GRAD_TOL = 2e-2, 3e-2      # bf16 operands (8 mantissa bits) at every processor GEMM / attention input


@pytest.mark.parametrize("name", ["vx2d", "dot2d", "pnet2d", "cond2d", "npe2d", "plain2d", "vxmix", "konly2d", "rope2d"])
def test_model_bf16_mode_option_cases(name):
    from gaot_b200 import ops
    from tests.helpers import golden_graphs, graphs_to
    c = cases.CASES[name]
    g = load_golden(name)
    ops.set_precision("bf16")
    try:
        model = build_model(c)
        model.load_state_dict({k: v.to(DEV) for k, v in golden_params(g).items()})
        lat, x, pn, tgt = [t.to(DEV) for t in cases.case_inputs(name)]
        enc = dec = None
        if c.get("vx"):
            enc, dec = graphs_to(golden_graphs(g, "enc", True), DEV), graphs_to(golden_graphs(g, "dec", True), DEV)
        cond = cases.case_condition(name)
        out = model(lat, x, pn, encoder_nbrs=enc, decoder_nbrs=dec, condition=cond.to(DEV) if cond is not None else None)
        loss = ops.mse_loss(out, tgt)
        loss.backward()
    finally:
        ops.set_precision("fp32")
    assert rel_l2(out, g["out"]) < BF16_FWD_TOL
    mine = torch.cat([p.grad.flatten().double().cpu() for _, p in model.named_parameters()])
    ref = torch.cat([torch.from_numpy(g["grad." + k]).flatten().double() for k, _ in model.named_parameters()])
    assert rel_l2(mine, ref) < BF16_GRAD_TOL


def test_model_bf16_mode_cfg2_geometry():
    """The benchmarked mode at the benchmarked geometry (BASELINE cfg-2: N=9216, latent 64x64, T=1024, Dh=256, 8 heads of 32),
    B=2, against the reference's fp64 summary fixture."""
    from gaot_b200 import ops
    from tests.helpers import golden_shapes
    c = cases.CASES["cfg2"]
    g = load_golden("cfg2")
    ops.set_precision("bf16")
    try:
        model = build_model(c)
        params = cases.synth_params(golden_shapes(g), 7 + c["seed"])
        model.load_state_dict({k: v.to(DEV) for k, v in params.items()})
        lat, x, pn, tgt = [t.to(DEV) for t in cases.case_inputs("cfg2")]
        inter = {}
        model.processor.register_forward_hook(lambda m, i, o: inter.__setitem__("proc", o.detach()))
        out = model(lat, x, pn)
        loss = ops.mse_loss(out, tgt)
        loss.backward()
    finally:
        ops.set_precision("fp32")
    assert rel_l2(inter["proc"][:, ::13, ::7], g["proc_sub"]) < BF16_FWD_TOL
    assert rel_l2(out[:, ::37, :], g["out_sub"]) < BF16_FWD_TOL
    assert abs(float(loss) - float(g["loss64"])) < 1e-2 * float(g["loss64"])
    grads = dict(model.named_parameters())
    norms = {str(k): float(n) for k, n in zip(g["keys"], g["grad_norms"])}
    tot_ref = sum(v * v for v in norms.values()) ** 0.5
    tot = sum(float(grads[k].grad.double().norm()) ** 2 for k in norms) ** 0.5
    assert abs(tot - tot_ref) < BF16_GRAD_TOL * tot_ref
    for k in [kk[5:] for kk in g if kk.startswith("grad.")]:
        assert rel_l2(grads[k].grad, g["grad." + k]) < 2 * BF16_GRAD_TOL, k


@pytest.mark.parametrize("use_graph", [False, True])
def test_rollout_bf16_mode_vs_reference_golden(use_graph):
    """The rollout fixture (4 steps, time_der) in bf16 mode: errors compound over the autoregressive steps."""
    from gaot_b200 import ops
    c, x0, tv, stats = cases.rollout_case()
    g = load_golden("rollout")
    ops.set_precision("bf16")
    try:
        model = build_model(c).eval()
        model.load_state_dict({k: v.to(DEV) for k, v in golden_params(g).items()})
        lat, x, _, _ = cases.case_inputs(cases.ROLLOUT["base"])
        got = model.autoregressive_predict(x0.to(DEV), cases.ROLLOUT["time_indices"], tv, stats, "time_der",
                                           latent_tokens_coord=lat.to(DEV), fixed_coord=x.to(DEV), use_cuda_graph=use_graph)
    finally:
        ops.set_precision("fp32")
    want = torch.from_numpy(g["pred.time_der"])
    for i in range(want.shape[1]):
        assert rel_l2(got[:, i], want[:, i]) < BF16_FWD_TOL, i


# ------------------------------------------------------------------------------------------------
# fused edge-weight kernels (gather + kernel MLP + attention weight, forward and backward, bf16 tensor cores)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,n_hidden,C,with_alpha", [("rand2d", 2, 64, True), ("rand3d", 1, 48, True), ("dec2d", 2, 16, False),
                                                        ("lattice2d", 1, 32, True), ("empty", 2, 64, True)])
def test_edge_mlp_fused_vs_fp64(name, n_hidden, C, with_alpha):
    """gaot_edge_mlp_{fwd,bwd} against an fp64 autograd restatement of agno.py:206-207, 229-242, 250.  Tolerances are those of
    the bf16 mode: hidden activations and hidden / output weights are rounded to bf16 (8 mantissa bits) at every MMA input,
    the first layer and all accumulations are fp32."""
    from oracle import gaot_oracle as orc
    from gaot_b200 import ops
    data, q, r = cases.nbr_case(name)
    g = orc.radius_csr(data, q, r)
    idx, rs = g["neighbors_index"], g["neighbors_row_splits"]
    kc, E = data.shape[1], int(idx.numel())
    plan = ops.GraphPlan(idx.to(DEV), rs.to(DEV), data.shape[0])
    gen = torch.Generator().manual_seed(11)
    dims = [2 * kc] + [64] * (n_hidden + 1) + [C]
    params = []
    for a, b in zip(dims[:-1], dims[1:]):
        params += [torch.randn(b, a, generator=gen) / a ** 0.5, 0.1 * torch.randn(b, generator=gen)]
    alpha = torch.rand(E, generator=gen) if with_alpha else None
    gw = torch.randn(E, C, generator=gen)
    pd = [p.clone().to(DEV).requires_grad_(True) for p in params]
    w = ops.edge_mlp(data.to(DEV), q.to(DEV), plan, alpha.to(DEV) if alpha is not None else None, pd)
    assert w.shape == (E, C)
    if E == 0:
        return
    w.backward(gw.to(DEV))
    pr = [p.clone().double().requires_grad_(True) for p in params]
    row = torch.repeat_interleave(torch.arange(q.shape[0]), rs[1:] - rs[:-1])
    h = torch.cat([data[idx], q[row]], 1).double()
    for i in range(0, len(pr) - 2, 2):
        h = torch.nn.functional.gelu(h @ pr[i].T + pr[i + 1])
    want = h @ pr[-2].T + pr[-1]
    if alpha is not None:
        want = want * alpha.double()[:, None]
    want.backward(gw.double())
    assert rel_l2(w, want) < 1e-2
    mine = torch.cat([p.grad.flatten().double().cpu() for p in pd])
    ref = torch.cat([p.grad.flatten() for p in pr])
    assert rel_l2(mine, ref) < 2e-2
    for a, b in zip(pd, pr):
        assert rel_l2(a.grad, b.grad) < 4e-2


def test_edge_mlp_fused_many_tiles_matches_unfused_bf16_chain():
    """E = 31 k edges (245 tiles, persistent CTAs take several tiles each: TMEM weight-gradient accumulators live across tiles)
    against the unfused bf16-mode chain on the same graph."""
    from gaot_b200 import ops
    from gaot_b200.model.layers.agno import AGNO
    x = cases.uniform_points(9216, 2, 0).to(DEV)
    lat = cases.latent_grid((64, 64)).to(DEV)
    plan = ops.graph_plan(ops.radius_csr(x, lat, 0.033), 9216)
    torch.manual_seed(3)
    agno = AGNO(channel_mlp_layers=[4, 64, 64, 64, 64], use_attn=True, attention_type="cosine", coord_dim=2).to(DEV)
    gw = cases.normal((plan.E, 64), 9).to(DEV)
    res = {}
    ops.set_precision("bf16")
    try:
        for fused in (True, False):
            agno.fuse_edge_mlp = fused
            agno.zero_grad()
            w = agno.edge_weights(x, lat, plan)
            w.backward(gw)
            res[fused] = (w.detach().clone(), torch.cat([p.grad.flatten() for p in agno.parameters()]))
    finally:
        ops.set_precision("fp32")
    assert rel_l2(res[True][0], res[False][0]) < 1e-2
    assert rel_l2(res[True][1], res[False][1]) < 2e-2


# ------------------------------------------------------------------------------------------------
# gated FFN with the SwiGLU fused into the GEMM epilogues (gaot_gemm_bf16_swiglu / gaot_gemm_bf16_nn_swiglu_bwd)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,D,F", [(256, 64, 128), (1000, 256, 1024), (4100, 64, 160), (32768, 256, 1024)])
def test_swiglu_ffn_fused_vs_fp64(M, D, F):
    """y = (silu(x w1^T) * (x w3^T)) w2^T + residual and its gradients against fp64 on the bf16-rounded operands
    (attn.py:137-146).  Tolerances: the saved pre-activations, u, dA | dB and dX are bf16 tensors (8 mantissa bits each):
    2e-2 on y and every gradient.  Shapes cover full and clipped tiles of the 256- and 128-column kernels."""
    from gaot_b200 import ops
    gen = torch.Generator().manual_seed(M + F)
    x = torch.randn(M, D, generator=gen).to(DEV).to(torch.bfloat16).requires_grad_(True)
    w1 = (torch.randn(F, D, generator=gen) / D ** 0.5).to(DEV).requires_grad_(True)
    w3 = (torch.randn(F, D, generator=gen) / D ** 0.5).to(DEV).requires_grad_(True)
    w2 = (torch.randn(D, F, generator=gen) / F ** 0.5).to(DEV).requires_grad_(True)
    res = torch.randn(M, D, generator=gen).to(DEV).requires_grad_(True)
    gy = torch.randn(M, D, generator=gen).to(DEV)
    ops.set_precision("bf16")
    try:
        assert ops.ffn_fusable(D, F)
        y = ops.swiglu_ffn(x, ops.interleave_w13(w1, w3), w2, add=res)
        y.backward(gy)
    finally:
        ops.set_precision("fp32")
    bf = lambda t: t.detach().to(torch.bfloat16).double()          # noqa: E731
    xr, w1r, w3r, w2r, rr = [bf(t).requires_grad_(True) for t in (x, w1, w3, w2)] + [res.detach().double().requires_grad_(True)]
    yr = (torch.nn.functional.silu(xr @ w1r.T) * (xr @ w3r.T)) @ w2r.T + rr
    yr.backward(bf(gy))
    assert rel_l2(y, yr) < 2e-2
    for got, ref, nm in ((x.grad, xr.grad, "dx"), (w1.grad, w1r.grad, "dw1"), (w3.grad, w3r.grad, "dw3"), (w2.grad, w2r.grad, "dw2"),
                         (res.grad, rr.grad, "dres")):
        assert rel_l2(got.float(), ref) < 2e-2, (nm, rel_l2(got.float(), ref))
