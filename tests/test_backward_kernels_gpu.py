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


# ------------------------------------------------------------------------------------------------ pointwise backward kernels
@pytest.mark.parametrize("rows,bias,gather", [(1000, False, False), (8 * 64, True, False), (3 * 64, False, True), (10368, False, False)])
def test_layernorm_bwd(rows, bias, gather):
    C = 768
    g = _gen(rows)
    if gather:   # rows 4..68 of each 324-token frame (Pcd_motion.py:520)
        Fr, L, M = rows // 64, 324, 64
        xfull = torch.randn(Fr * L, C, generator=g).to(DEV) * 3 + 0.5
        idx = (torch.arange(rows) // M) * L + 4 + torch.arange(rows) % M
        x = xfull[idx.to(DEV)]
    else:
        xfull = x = (torch.randn(rows, C, generator=g) * 3 + 0.5).to(DEV)
    w = (1 + 0.2 * torch.randn(C, generator=g)).to(DEV)
    b = (0.1 * torch.randn(C, generator=g)).to(DEV) if bias else None
    dy = (torch.randn(rows, C, generator=g) * 1e-2).to(DEV)
    dres = None if gather else (torch.randn(rows, C, generator=g) * 1e-2).to(DEV)
    alpha = 2.5e-3
    xd = x.double().requires_grad_(True); wd = w.double().requires_grad_(True)
    bd = b.double().requires_grad_(True) if bias else None
    torch.nn.functional.layer_norm(xd, (C,), wd, bd, 1e-5).backward(dy.double())
    dgamma = torch.full((C,), 1e-3, device=DEV)     # accumulates (+=) like .grad
    dbeta = torch.full((C,), -1e-3, device=DEV) if bias else None
    if gather:
        dx32 = torch.zeros(Fr * L, C, device=DEV)
        ops.layernorm_bwd(dy, xfull, w, 1e-5, rows, C, src_rpg=M, src_gstride=L, src_goff=4, dx32=dx32, lddx32=C, dgamma=dgamma, alpha=alpha)
        got = dx32[idx.to(DEV)]
        mask = torch.ones(Fr * L, dtype=torch.bool); mask[idx] = False
        assert float(dx32[mask.to(DEV)].abs().max()) == 0.0          # only the gathered rows are written
        ref_dx = xd.grad
    else:
        dx32 = dres.clone()
        dx16 = torch.empty(rows, C, device=DEV, dtype=torch.float16)
        ops.layernorm_bwd(dy, x, w, 1e-5, rows, C, dres=dx32, lddres=C, dx32=dx32, lddx32=C, dx16=dx16, lddx16=C, dgamma=dgamma,
                          dbeta=dbeta, alpha=alpha)
        got, ref_dx = dx32, xd.grad + dres.double()
        assert _rel(dx16, ref_dx) < 1e-3
    assert _rel(got, ref_dx) < 2e-6, _rel(got, ref_dx)
    assert _rel(dgamma - 1e-3, alpha * wd.grad) < 1e-4, _rel(dgamma - 1e-3, alpha * wd.grad)
    if bias:
        assert _rel(dbeta + 1e-3, alpha * bd.grad) < 1e-4


def test_qknorm_bwd():
    rows, d = 700, 768
    g = _gen(21)
    raw = torch.randn(rows, 3 * d, generator=g).to(DEV)
    wq, wk = (1 + 0.2 * torch.randn(64, generator=g)).to(DEV), (1 + 0.2 * torch.randn(64, generator=g)).to(DEV)
    d_in = (torch.randn(rows, 3 * d, generator=g) * 1e-2).to(DEV)
    rd = raw.double().requires_grad_(True); wqd = wq.double().requires_grad_(True); wkd = wk.double().requires_grad_(True)
    q, k, v = rd[:, :d].reshape(rows, 12, 64), rd[:, d:2 * d].reshape(rows, 12, 64), rd[:, 2 * d:]
    rq, rk = torch.rsqrt((q * q).mean(-1, keepdim=True) + 1e-5), torch.rsqrt((k * k).mean(-1, keepdim=True) + 1e-5)
    y = torch.cat([(q * rq * wqd).reshape(rows, d), (k * rk * wkd).reshape(rows, d), v], dim=1)
    y.backward(d_in.double())
    y16 = y.detach().half()
    rstd = torch.cat([rq.detach().reshape(rows, 12), rk.detach().reshape(rows, 12)], dim=1).float().contiguous()
    out = torch.empty(rows, 3 * d, device=DEV, dtype=torch.float16)
    dwq, dwk = torch.zeros(64, device=DEV), torch.zeros(64, device=DEV)
    alpha = 0.01
    ops.qknorm_bwd(d_in, 3 * d, y16, 3 * d, rstd, 24, wq, wk, d, 2 * d, 3 * d, rows, out, 3 * d, dwq, dwk, alpha)
    assert _rel(out, rd.grad) < 1.5e-3, _rel(out, rd.grad)        # fp16 y -> xhat and fp16 output
    assert _rel(dwq, alpha * wqd.grad) < 1e-3 and _rel(dwk, alpha * wkd.grad) < 1e-3


def test_head_bwd_and_colsum():
    rows, C = 5000, 768
    g = _gen(33)
    u = torch.randn(rows, C, generator=g).to(DEV)
    w3 = (torch.randn(3, C, generator=g) * 0.02).to(DEV)
    b3 = torch.zeros(3)
    target = torch.randn(rows, 3, generator=g).to(DEV) * 0.3
    ud = u.double().requires_grad_(True); w3d = w3.double().requires_grad_(True); b3d = b3.double().to(DEV).requires_grad_(True)
    pred = torch.nn.functional.gelu(ud) @ w3d.t() + b3d
    loss_unscaled = 0.5 * ((pred - target.double()) ** 2).sum()      # d/dpred = pred - target: the seed in units of 1/alpha
    loss_unscaled.backward()
    du = torch.empty(rows, C, device=DEV, dtype=torch.float16)
    dw3, db3 = torch.zeros(3, C, device=DEV), torch.zeros(3, device=DEV)
    alpha = 1e-4
    ops.head_bwd(pred.detach().float().contiguous(), target, u, C, w3, rows, C, du, C, dw3, db3, alpha)
    assert _rel(du, ud.grad) < 1e-3, _rel(du, ud.grad)
    assert _rel(dw3, alpha * w3d.grad) < 1e-4 and _rel(db3, alpha * b3d.grad) < 1e-4
    db = torch.zeros(C, device=DEV)
    ops.colsum(du, C, rows, C, db, alpha)
    assert _rel(db, alpha * du.double().sum(0)) < 1e-5


def test_sum_groups_transpose_dot():
    g = _gen(4)
    T, N, C = 7, 300, 768
    x = torch.randn(T * N, C, generator=g).to(DEV)
    out32 = torch.ones(N, C, device=DEV)
    out16 = torch.empty(N, C, device=DEV, dtype=torch.float16)
    ops.sum_groups(x, C, T, N, N, C, scale=0.5, accumulate=1, out32=out32, ldo32=C, out16=out16, ldo16=C)
    ref = 1 + 0.5 * x.double().reshape(T, N, C).sum(0)
    assert _rel(out32, ref) < 1e-6 and _rel(out16, ref) < 1e-3
    # mesh-token rows 4..68 of every frame, summed over the frames of one clip
    L, M = 324, 64
    tok = torch.randn(T * L, C, generator=g).to(DEV)
    dm = torch.empty(M, C, device=DEV)
    ops.sum_groups(tok, C, T, L, M, C, rpg=M, in_gstride=0, in_goff=4, out32=dm, ldo32=C)
    assert _rel(dm, tok.double().reshape(T, L, C)[:, 4:68].sum(0)) < 1e-6
    W = torch.randn(770, 100, generator=g).to(DEV)
    Wt = torch.full((100, 832), 7.0, device=DEV, dtype=torch.float16)
    ops.cast_transpose_f16(W, 770, 100, Wt, 832, npad=832)
    assert torch.equal(Wt[:, :770], W.t().half()) and float(Wt[:, 770:].abs().max()) == 0.0
    rows, H = 1000, 12
    dO, O = torch.randn(rows, 768, generator=g).to(DEV).half(), torch.randn(rows, 768, generator=g).to(DEV).half()
    D = torch.empty(rows, H, device=DEV)
    ops.attn_dot(dO, 768, O, 768, rows, H, D, H)
    assert _rel(D, (dO.double() * O.double()).reshape(rows, H, 64).sum(-1)) < 1e-6


# ------------------------------------------------------------------------------------------------ attention backward
def _attn_ref(q, k, v, dO, scale):
    """fp64 autograd of softmax(q k^T * scale) v per (batch, head); q [B,Lq,H,64], k/v [B,Lk,H,64]."""
    qd, kd, vd = (t.double().requires_grad_(True) for t in (q, k, v))
    s = torch.einsum("bqhd,bkhd->bhqk", qd, kd) * scale
    o = torch.einsum("bhqk,bkhd->bqhd", torch.softmax(s, dim=-1), vd)
    o.backward(dO.double())
    lse2 = torch.logsumexp(s, dim=-1) * 1.4426950408889634      # [B,H,Lq], log2 domain
    return o.detach(), qd.grad, kd.grad, vd.grad, lse2.permute(0, 2, 1)


@pytest.mark.parametrize("B,Lq,Lk,H", [(1, 128, 128, 1), (2, 256, 384, 2), (1, 324, 324, 12), (3, 200, 64, 2), (1, 1000, 777, 3)])
def test_attention_lse_and_backward_packed_qkv(B, Lq, Lk, H):
    d = H * 64
    g = _gen(B * 1000 + Lq + Lk)
    q = torch.randn(B, Lq, H, 64, generator=g).to(DEV).half()
    k = torch.randn(B, Lk, H, 64, generator=g).to(DEV).half()
    v = torch.randn(B, Lk, H, 64, generator=g).to(DEV).half()
    dO = (torch.randn(B, Lq, H, 64, generator=g) * 1e-2).to(DEV).half()
    scale = 0.125
    o_ref, dq_ref, dk_ref, dv_ref, lse_ref = _attn_ref(q, k, v, dO, scale)
    q2, k2, v2, dO2 = (t.reshape(-1, d).contiguous() for t in (q, k, v, dO))
    out = torch.empty(B * Lq, d, device=DEV, dtype=torch.float16)
    lse = torch.full((B * Lq, H), float("nan"), device=DEV)
    ops.attention(q2, k2, v2, out, B=B, H=H, Lq=Lq, Lk=Lk, q_ld=d, k_ld=d, v_ld=d, o_ld=d, q_rows=B * Lq, kv_rows=B * Lk,
                  q_batch_rows=Lq, kv_batch_rows=Lk, scale=scale, lse=lse, lse_ld=H)
    assert _rel(out, o_ref.reshape(-1, d)) < 2e-3
    assert float((lse.double() - lse_ref.reshape(-1, H)).abs().max()) < 2e-3     # log2 units
    D = torch.empty(B * Lq, H, device=DEV)
    ops.attn_dot(dO2, d, out, d, B * Lq, H, D, H)
    dQ = torch.zeros(B * Lq, d, device=DEV)
    dK = torch.full((B * Lk, d), float("nan"), device=DEV)
    dV = torch.full((B * Lk, d), float("nan"), device=DEV)
    ops.attention_bwd(q2, k2, v2, dO2, lse, D, dQ, dK, dV, B=B, H=H, Lq=Lq, Lk=Lk, q_ld=d, k_ld=d, v_ld=d, do_ld=d, lse_ld=H, d_ld=H,
                      dq_ld=d, dk_ld=d, dv_ld=d, q_rows=B * Lq, kv_rows=B * Lk, q_batch_rows=Lq, kv_batch_rows=Lk, scale=scale)
    assert torch.isfinite(dK).all() and torch.isfinite(dV).all()
    # fp16 P / dS operands (2^-11 per element) and fp16 O inside D
    assert _rel(dV, dv_ref.reshape(-1, d)) < 2e-3, _rel(dV, dv_ref.reshape(-1, d))
    assert _rel(dQ, dq_ref.reshape(-1, d)) < 3e-3, _rel(dQ, dq_ref.reshape(-1, d))
    assert _rel(dK, dk_ref.reshape(-1, d)) < 3e-3, _rel(dK, dk_ref.reshape(-1, d))


def test_attention_backward_shared_query_accumulates_over_batches():
    """The decoder's addressing (Pcd_motion.py:539-560): one query operand for all frames, K/V per frame."""
    T, N, M, H = 3, 300, 64, 2
    d = H * 64
    g = _gen(77)
    q = torch.randn(1, N, H, 64, generator=g).to(DEV).half()
    k = torch.randn(T, M, H, 64, generator=g).to(DEV).half()
    v = torch.randn(T, M, H, 64, generator=g).to(DEV).half()
    dO = (torch.randn(T, N, H, 64, generator=g) * 1e-2).to(DEV).half()
    o_ref, dq_ref, dk_ref, dv_ref, _ = _attn_ref(q.expand(T, -1, -1, -1).contiguous(), k, v, dO, 0.125)
    kv = torch.cat([k.reshape(-1, d), v.reshape(-1, d)], dim=1).contiguous()     # packed k | v rows
    out = torch.empty(T * N, d, device=DEV, dtype=torch.float16)
    lse = torch.empty(T * N, H, device=DEV)
    common = dict(B=T, H=H, Lq=N, Lk=M, q_ld=d, k_ld=2 * d, v_ld=2 * d, q_rows=N, kv_rows=T * M, q_batch_rows=0, kv_batch_rows=M, scale=0.125)
    ops.attention(q.reshape(-1, d), kv, kv[:, d:], out, o_ld=d, lse=lse, lse_ld=H, **common)
    D = torch.empty(T * N, H, device=DEV)
    ops.attn_dot(dO.reshape(-1, d), d, out, d, T * N, H, D, H)
    dQ = torch.zeros(N, d, device=DEV)
    dKV = torch.zeros(T * M, 2 * d, device=DEV)
    ops.attention_bwd(q.reshape(-1, d), kv, kv[:, d:], dO.reshape(-1, d), lse, D, dQ, dKV, dKV[:, d:], do_ld=d, lse_ld=H, d_ld=H, dq_ld=d,
                      dk_ld=2 * d, dv_ld=2 * d, **common)
    assert _rel(dQ, dq_ref.sum(0).reshape(-1, d)) < 3e-3
    assert _rel(dKV[:, :d], dk_ref.reshape(-1, d)) < 3e-3 and _rel(dKV[:, d:], dv_ref.reshape(-1, d)) < 2e-3
