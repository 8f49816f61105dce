"""GPU parity tests of the backward-pass kernels through the C ABI against fp64 torch autograd /
closed-form restatements of the same operator.  Gradient operands travel as fp16 in units of 1/alpha (alpha = dLoss*2w/n,
so the seed is pred - target), saved activations and weights as fp16, all accumulation in fp32; tolerances below are
relative L2 against fp64 on the SAME rounded operands (tight: only accumulation order differs) or on unrounded ones
(fp16 rounding of an output: 2^-11 per element)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from motion324_b200 import ops  # noqa: E402

DEV = "cuda"


def _rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _gen(seed):
    return torch.Generator(device="cpu").manual_seed(seed)


# ------------------------------------------------------------------------------------------------ GEMM: dgrad / wgrad
@pytest.mark.parametrize("M,N,K", [(256, 768, 768), (1000, 768, 3072), (10368, 3072, 768), (64, 768, 2304), (4100, 768, 1536)])
def test_gemm_dgrad(M, N, K):
    """dX[M, K_in=N] = dY[M, K=N_out] . W  with A = dY and the transposed weight W^T [K_in, N_out]; also all-bf16 operands."""
    g = _gen(M + N + K)
    for dt, tol16 in ((torch.float16, 1e-3), (torch.bfloat16, 4e-3)):
        dY = (torch.randn(M, K, generator=g) * 1e-2).to(DEV).to(dt)
        Wt = (torch.randn(N, K, generator=g) * 0.05).to(DEV).to(dt)
        out32 = torch.full((M, N), float("nan"), device=DEV)
        out16 = torch.zeros(M, N, device=DEV, dtype=dt)
        ops.gemm(dY, Wt, M, N, K, out32=out32, ldo32=N, out16=out16, ldo16=N)
        ref = dY.double() @ Wt.double().t()
        assert _rel(out32, ref) < 1e-5, _rel(out32, ref)
        assert _rel(out16, ref) < tol16


@pytest.mark.parametrize("rows,Nout,Kin,ksplit", [(64, 768, 768, 1), (4096, 768, 768, 8), (10368, 2304, 768, 6), (10368, 768, 3072, 4),
                                                  (1000, 768, 832, 3), (333, 1536, 64, 2), (131072, 768, 768, 32)])
def test_gemm_wgrad_tn_splitk_accumulate(rows, Nout, Kin, ksplit):
    """dW[N_out, K_in] += alpha * dY^T . X straight from row-major dY [rows, N_out] and X [rows, K_in] (fp16)."""
    g = _gen(rows + Nout + Kin)
    dY = (torch.randn(rows, Nout, generator=g) * 1e-2).to(DEV).half()
    X = torch.randn(rows, Kin, generator=g).to(DEV).half()
    prev = (torch.randn(Nout, Kin, generator=g) * 1e-4).to(DEV)
    dW = prev.clone()
    alpha = 3.0e-3
    ops.gemm(dY, X, Nout, Kin, rows, lda=Nout, ldw=Kin, tn=1, ksplit=ksplit, accumulate=1, out32=dW, ldo32=Kin, out_scale=alpha)
    ref = prev.double() + alpha * (dY.double().t() @ X.double())
    assert torch.isfinite(dW).all()
    assert _rel(dW, ref) < 2e-5, _rel(dW, ref)


def test_gemm_aux_pre_activation_and_gelu_backward():
    """Forward (training): out16 = gelu(u), aux16 = u.  Backward: dU = (dG . W) * gelu'(u) in the epilogue."""
    M, N, K = 700, 3072, 768
    g = _gen(5)
    A = torch.randn(M, K, generator=g).to(DEV).half()
    W = (torch.randn(N, K, generator=g) * 0.05).to(DEV).half()
    hid = torch.empty(M, N, device=DEV, dtype=torch.float16)
    pre = torch.empty(M, N, device=DEV, dtype=torch.float16)
    ops.gemm(A, W, M, N, K, act=1, out16=hid, ldo16=N, aux16=pre, ldaux=N, aux_mode=1)
    u = A.double() @ W.double().t()
    assert _rel(pre, u) < 1e-3 and _rel(hid, torch.nn.functional.gelu(u)) < 1e-3
    # backward: dG [M, 768] . W2t [3072, 768]^T -> [M, 3072], times gelu'(pre)
    dG = (torch.randn(M, K, generator=g) * 1e-2).to(DEV).half()
    W2t = (torch.randn(N, K, generator=g) * 0.05).to(DEV).half()
    dU = torch.empty(M, N, device=DEV, dtype=torch.float16)
    ops.gemm(dG, W2t, M, N, K, out16=dU, ldo16=N, aux16=pre, ldaux=N, aux_mode=2)
    up = pre.double().requires_grad_(True)
    torch.nn.functional.gelu(up).backward(dG.double() @ W2t.double().t())
    assert _rel(dU, up.grad) < 1e-3, _rel(dU, up.grad)


def test_gemm_qk_rstd_output():
    M, d = 500, 768
    g = _gen(9)
    A = torch.randn(M, d, generator=g).to(DEV).half()
    W = (torch.randn(3 * d, d, generator=g) * 0.05).to(DEV).half()
    qn, kn = (1 + 0.1 * torch.randn(64, generator=g)).to(DEV), (1 + 0.1 * torch.randn(64, generator=g)).to(DEV)
    out = torch.empty(M, 3 * d, device=DEV, dtype=torch.float16)
    rstd = torch.full((M, 24), float("nan"), device=DEV)
    ops.gemm(A, W, M, 3 * d, d, out16=out, ldo16=3 * d, qn_w=qn, kn_w=kn, qk_eps=1e-5, qk_cols=d, qk_rstd=rstd, ld_rstd=24)
    raw = (A.double() @ W.double().t())[:, :2 * d].reshape(M, 24, 64)
    ref = torch.rsqrt((raw * raw).mean(-1) + 1e-5)
    assert _rel(rstd, ref) < 1e-5
