"""GPU parity tests of the individual libm324 kernels (called through the C ABI) against plain fp32/fp64 torch
restatements of the same op, and against the oracle's operator functions.  Run on the B200 box: pytest -m gpu."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():  # collected on CPU, skipped there
    pytest.skip("needs a CUDA device", allow_module_level=True)

from motion324_b200 import ops  # noqa: E402
from oracle import motion324_oracle as orc  # noqa: E402

DEV = "cuda"


def _rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _gen(seed):
    return torch.Generator(device="cpu").manual_seed(seed)


def test_device_is_sm100():
    ops.check_device()


# mode: 0 auto (2-CTA pair kernel when M > 128), 1 = 1-CTA 128x128, 2 = 1-CTA 128x256, 3 = 2-CTA 256x128, 4 = 2-CTA 256x256
@pytest.mark.parametrize("M,N,K,bn128", [(128, 256, 64, 2), (256, 256, 128, 2), (1000, 768, 768, 2), (300, 2304, 768, 2),
                                          (128, 128, 64, 1), (515, 768, 3072, 1), (64, 768, 768, 0), (4096, 3072, 768, 2),
                                          (256, 256, 64, 4), (512, 256, 128, 4), (1000, 768, 768, 4), (300, 2304, 768, 0),
                                          (515, 768, 3072, 3), (4096, 3072, 768, 0), (10368, 768, 3072, 0), (129, 128, 64, 3)])
def test_gemm_plain(M, N, K, bn128):
    g = _gen(M + N + K)
    A = (torch.randn(M, K, generator=g)).to(DEV).half()
    W = (torch.randn(N, K, generator=g) * 0.05).to(DEV).half()
    out32 = torch.full((M, N), float("nan"), device=DEV)
    out16 = torch.zeros(M, N, device=DEV, dtype=torch.float16)
    ops.gemm(A, W, M, N, K, out32=out32, ldo32=N, out16=out16, ldo16=N, force_bn128=bn128)
    ref = A.double() @ W.double().t()
    assert torch.isfinite(out32).all()
    assert _rel(out32, ref) < 6e-6, _rel(out32, ref)  # fp32 accumulation over K, fp16-exact operands
    assert _rel(out16, ref) < 1e-3


def test_gemm_epilogue_bias_gelu_gamma_resid():
    M, N, K = 777, 768, 768
    g = _gen(3)
    A = torch.randn(M, K, generator=g).to(DEV).half()
    W = (torch.randn(N, K, generator=g) * 0.05).to(DEV).half()
    bias = torch.randn(N, generator=g).to(DEV)
    gamma = (1 + 0.1 * torch.randn(N, generator=g)).to(DEV)
    resid = torch.randn(100, N, generator=g).to(DEV)
    out32 = torch.empty(M, N, device=DEV)
    ops.gemm(A, W, M, N, K, bias=bias, gamma=gamma, act=1, resid=resid, ldr=N, resid_mod=100, out32=out32, ldo32=N)
    y = torch.nn.functional.gelu(A.double() @ W.double().t() + bias.double()) * gamma.double()
    ref = y + resid.double()[torch.arange(M, device=DEV) % 100]
    assert _rel(out32, ref) < 3e-6, _rel(out32, ref)
    # resid_div mapping: row -> (row // 300) * 30 + row % 30
    out32b = torch.empty(M, N, device=DEV)
    ops.gemm(A, W, M, N, K, resid=resid, ldr=N, resid_mod=30, resid_div=300, out32=out32b, ldo32=N)
    rows = torch.arange(M, device=DEV)
    refb = A.double() @ W.double().t() + resid.double()[(rows // 300) * 30 + rows % 30]
    assert _rel(out32b, refb) < 3e-6
    # in-place residual (out32 aliases resid)
    x = torch.randn(M, N, generator=g).to(DEV)
    x0 = x.clone()
    ops.gemm(A, W, M, N, K, resid=x, ldr=N, out32=x, ldo32=N)
    assert _rel(x, A.double() @ W.double().t() + x0.double()) < 3e-6


def test_gemm_qk_norm_epilogue():
    M, K = 700, 768
    N = 2304
    g = _gen(4)
    A = torch.randn(M, K, generator=g).to(DEV).half()
    W = (torch.randn(N, K, generator=g) * 0.05).to(DEV).half()
    qw = (1 + 0.2 * torch.randn(64, generator=g)).to(DEV)
    kw = (1 + 0.2 * torch.randn(64, generator=g)).to(DEV)
    out16 = torch.empty(M, N, device=DEV, dtype=torch.float16)
    ops.gemm(A, W, M, N, K, out16=out16, ldo16=N, qn_w=qw, kn_w=kw, qk_eps=1e-5, qk_cols=768)
    y = (A.double() @ W.double().t())
    q, k, v = y[:, :768].reshape(M, 12, 64), y[:, 768:1536].reshape(M, 12, 64), y[:, 1536:]
    q = orc.rms_norm(q, qw.double()).reshape(M, 768)
    k = orc.rms_norm(k, kw.double()).reshape(M, 768)
    ref = torch.cat([q, k, v], dim=1)
    assert _rel(out16, ref) < 6e-4, _rel(out16, ref)
    # k-norm only on the first 768 columns of a [k | v] projection (cross attention)
    out16b = torch.empty(M, 1536, device=DEV, dtype=torch.float16)
    ops.gemm(A, W[768:], M, 1536, K, out16=out16b, ldo16=1536, qn_w=kw, kn_w=None, qk_cols=768)
    assert _rel(out16b, torch.cat([k, v], dim=1)) < 6e-4


def test_gemm_split_precision():
    M, N, K = 640, 768, 832
    g = _gen(5)
    A = torch.randn(M, K, generator=g).to(DEV)
    W = (torch.randn(N, K, generator=g) * 0.05).to(DEV)
    A16 = torch.empty(M, 2 * K, device=DEV, dtype=torch.float16)
    W16 = torch.empty(N, 2 * K, device=DEV, dtype=torch.float16)
    ops.cast_pad_f16(A, M, K, A16, 2 * K, K, lo_off=K)
    ops.cast_pad_f16(W, N, K, W16, 2 * K, K, lo_off=K)
    hi = A.half()
    assert torch.equal(A16[:, :K], hi) and torch.equal(A16[:, K:], (A - hi.float()).half())
    out32 = torch.empty(M, N, device=DEV)
    out16 = torch.empty(M, 2 * N, device=DEV, dtype=torch.float16)
    ops.gemm(A16, W16, M, N, K, passes=3, a_lo_off=K, w_lo_off=K, out32=out32, ldo32=N, out16=out16, ldo16=2 * N,
             out16_lo_off=N)
    ref = A.double() @ W.double().t()
    assert _rel(out32, ref) < 5e-6, _rel(out32, ref)
    assert _rel(out16[:, :N].float() + out16[:, N:].float(), ref) < 5e-6
    # the plain fp16 path on the same data is ~100x worse: proves the split does something
    o1 = torch.empty(M, N, device=DEV)
    ops.gemm(A16, W16, M, N, K, out32=o1, ldo32=N)
    assert _rel(o1, ref) > 1e-4


def _attn_ref(q, k, v, scale):
    # q [B,Lq,H,D] etc, fp64 math on the fp16-rounded operands
    q_, k_, v_ = (t.double().transpose(1, 2) for t in (q, k, v))
    s = (q_ @ k_.transpose(-2, -1)) * scale
    return (torch.softmax(s, dim=-1) @ v_).transpose(1, 2)


@pytest.fixture(params=[1, 2], ids=["pair-kernel", "split-kernel"])
def attn_mode(request):
    """Force the work-item shape of the attention kernel: 1 = two Q tiles sharing the K/V walk, 2 = one Q tile with the
    K/V range split between the two softmax groups and merged in the CTA (falls back to 1 for a single K/V tile), 3 = two
    threads per query row (16 softmax warps)."""
    ops.set_tuning(0, request.param)
    yield request.param
    ops.set_tuning(0, 0)


@pytest.mark.parametrize("B,H,Lq,Lk", [(1, 12, 256, 128), (2, 12, 324, 324), (3, 12, 257, 257), (1, 12, 64, 64),
                                        (1, 12, 64, 4096), (1, 12, 1296, 1296), (2, 3, 100, 700), (1, 2, 130, 1150)])
def test_attention_self_and_cross(B, H, Lq, Lk, attn_mode):
    g = _gen(B * 1000 + Lq + Lk)
    q = torch.randn(B, Lq, H, 64, generator=g).to(DEV).half()
    k = torch.randn(B, Lk, H, 64, generator=g).to(DEV).half()
    v = torch.randn(B, Lk, H, 64, generator=g).to(DEV).half()
    out = torch.zeros(B * Lq, H * 64, device=DEV, dtype=torch.float16)
    ops.attention(q, k, v, out, B=B, H=H, Lq=Lq, Lk=Lk, q_ld=H * 64, k_ld=H * 64, v_ld=H * 64, o_ld=H * 64,
                  q_rows=B * Lq, kv_rows=B * Lk, q_batch_rows=Lq, kv_batch_rows=Lk, scale=0.125)
    ref = _attn_ref(q, k, v, 0.125).reshape(B * Lq, H * 64)
    assert torch.isfinite(out).all()
    assert _rel(out, ref) < 1.5e-3, _rel(out, ref)


@pytest.mark.parametrize("B,H,Lq,Lk", [(1, 12, 10368, 10368), (32, 12, 324, 324), (32, 12, 257, 257), (32, 12, 4096, 64)])
def test_attention_at_the_benchmarked_shapes(B, H, Lq, Lk):
    """The launches of BASELINE config (b) as they run in bench.py (packed [rows, 3*768] q|k|v of a trunk block, module
    workspace -> tail split + merge for the 10368-token global layer; 324 / 257 tokens per frame for the local and DINOv2
    blocks; the decoder's 4096 shared queries x 64 keys per frame) against fp64 softmax attention, one head at a time."""
    g = _gen(Lq + Lk)
    d = H * 64
    shared_q = Lk == 64
    qkv = torch.randn(B * Lk, 3 * d, generator=g).to(DEV).half()
    q = torch.randn(Lq, d, generator=g).to(DEV).half() if shared_q else qkv
    out = torch.zeros(B * Lq, d, device=DEV, dtype=torch.float16)
    ops.attention(q, qkv[:, d:], qkv[:, 2 * d:], out, B=B, H=H, Lq=Lq, Lk=Lk, q_ld=d if shared_q else 3 * d, k_ld=3 * d, v_ld=3 * d,
                  o_ld=d, q_rows=Lq if shared_q else B * Lq, kv_rows=B * Lk, q_batch_rows=0 if shared_q else Lq, kv_batch_rows=Lk, scale=0.125)
    assert torch.isfinite(out).all()
    num = den = 0.0
    for h in range(H):
        qh = (q[:, h * 64:(h + 1) * 64].double().view(1, Lq, 64).expand(B, Lq, 64) if shared_q
              else qkv[:, h * 64:(h + 1) * 64].double().view(B, Lq, 64))
        kh = qkv[:, d + h * 64:d + (h + 1) * 64].double().view(B, Lk, 64)
        vh = qkv[:, 2 * d + h * 64:2 * d + (h + 1) * 64].double().view(B, Lk, 64)
        ref = torch.softmax(qh @ kh.transpose(1, 2) * 0.125, dim=-1) @ vh
        got = out[:, h * 64:(h + 1) * 64].double().view(B, Lq, 64)
        num += float((got - ref).pow(2).sum())
        den += float(ref.pow(2).sum())
    rel = (num / den) ** 0.5
    assert rel < 1.5e-3, rel


@pytest.mark.parametrize("B,H,Lq,Lk", [(32, 12, 324, 324), (40, 12, 64, 64), (30, 12, 300, 500), (50, 6, 129, 257), (26, 12, 256, 128), (13, 12, 700, 40)])
def test_attention_item_loop_matches_one_cta_per_item(B, H, Lq, Lk):
    """attn_items_kernel (persistent CTAs over contiguous chunks of short work items: local / DINOv2 / latent blocks) against fp64
    softmax attention AND against the one-CTA-per-item kernel (knob 1 = 2 switches the loop off): outputs and log-sum-exp.
    Covers a ragged second Q tile (group 0 alone), a Q tile pair with an out-of-range second tile, 1 .. 4 K/V tiles per item,
    partial last K/V tiles, and chunk boundaries that fall inside a (batch, head)."""
    import math
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    assert ((Lq + 255) // 256) * H * B >= 2 * sms and (Lk + 127) // 128 <= 4      # the shapes above do take the item loop
    g = _gen(7 * B + Lq + Lk)
    q = torch.randn(B, Lq, H, 64, generator=g).to(DEV).half()
    k = (torch.randn(B, Lk, H, 64, generator=g) * torch.linspace(0.5, 2.0, Lk).view(1, Lk, 1, 1)).to(DEV).half()
    v = torch.randn(B, Lk, H, 64, generator=g).to(DEV).half()
    kw = dict(B=B, H=H, Lq=Lq, Lk=Lk, q_ld=H * 64, k_ld=H * 64, v_ld=H * 64, o_ld=H * 64, q_rows=B * Lq, kv_rows=B * Lk,
              q_batch_rows=Lq, kv_batch_rows=Lk, scale=0.125, lse_ld=H)
    out, out0 = (torch.zeros(B * Lq, H * 64, device=DEV, dtype=torch.float16) for _ in range(2))
    lse, lse0 = (torch.zeros(B * Lq, H, device=DEV) for _ in range(2))
    ops.attention(q, k, v, out, lse=lse, **kw)
    ops.set_tuning(1, 2)
    try:
        ops.attention(q, k, v, out0, lse=lse0, **kw)
    finally:
        ops.set_tuning(1, 0)
    torch.cuda.synchronize()
    ref = _attn_ref(q, k, v, 0.125).reshape(B * Lq, H * 64)
    assert torch.isfinite(out).all()
    assert _rel(out, ref) < 1.5e-3 and _rel(out0, ref) < 1.5e-3, (_rel(out, ref), _rel(out0, ref))
    assert torch.equal(out, out0) and torch.equal(lse, lse0)       # same tile order per item -> bit-identical to the per-item kernel
    s = (q.double().transpose(1, 2) @ k.double().transpose(1, 2).transpose(-2, -1)) * 0.125
    lse_ref = (torch.logsumexp(s, dim=-1) / math.log(2.0)).transpose(1, 2).reshape(B * Lq, H)
    assert float((lse.double() - lse_ref).abs().max()) < 2e-3


@pytest.mark.parametrize("B,H,Lq,Lk,parts", [(1, 12, 3300, 3300, 4), (2, 12, 1700, 1100, 2), (13, 12, 200, 2100, 4)])
def test_attention_tail_split_of_the_last_wave(B, H, Lq, Lk, parts):
    """Work items of a partly filled last wave are split over K/V ranges (workspace given) and merged by a second kernel:
    same result as the unsplit launch, log-sum-exp included; ragged last Q and K/V tiles inside a split item."""
    import math
    items = ((Lq + 255) // 256) * H * B
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    rem = items % sms
    assert items > sms and 0 < rem <= sms // 2 and min(sms // rem, 4, ((Lk + 127) // 128) // 4) == parts   # the shapes above do split
    g = _gen(B + Lq + Lk)
    q = torch.randn(B, Lq, H, 64, generator=g).to(DEV).half()
    k = (torch.randn(B, Lk, H, 64, generator=g) * torch.linspace(0.5, 2.0, Lk).view(1, Lk, 1, 1)).to(DEV).half()   # later K/V ranges hold the maxima
    v = torch.randn(B, Lk, H, 64, generator=g).to(DEV).half()
    kw = dict(B=B, H=H, Lq=Lq, Lk=Lk, q_ld=H * 64, k_ld=H * 64, v_ld=H * 64, o_ld=H * 64, q_rows=B * Lq, kv_rows=B * Lk,
              q_batch_rows=Lq, kv_batch_rows=Lk, scale=0.125, lse_ld=H)
    out, out0 = (torch.zeros(B * Lq, H * 64, device=DEV, dtype=torch.float16) for _ in range(2))
    lse, lse0 = (torch.zeros(B * Lq, H, device=DEV) for _ in range(2))
    n0 = ops.LAUNCHES[0]
    ops.attention(q, k, v, out, lse=lse, **kw)                       # module workspace -> tail split
    ops.attention(q, k, v, out0, lse=lse0, workspace=None, **kw)     # no workspace -> every item walks all of K/V
    ref = _attn_ref(q, k, v, 0.125).reshape(B * Lq, H * 64)
    assert torch.isfinite(out).all()
    assert _rel(out, ref) < 1.5e-3 and _rel(out0, ref) < 1.5e-3, (_rel(out, ref), _rel(out0, ref))
    assert _rel(out, out0) < 1e-3
    s = (q.double().transpose(1, 2) @ k.double().transpose(1, 2).transpose(-2, -1)) * 0.125
    lse_ref = (torch.logsumexp(s, dim=-1) / math.log(2.0)).transpose(1, 2).reshape(B * Lq, H)
    assert float((lse.double() - lse_ref).abs().max()) < 2e-3 and float((lse0.double() - lse_ref).abs().max()) < 2e-3
    assert not torch.equal(out, out0)      # the split path really ran (different summation order)


@pytest.mark.parametrize("B,H,Lq,ranges", [(1, 12, 700, [(300, 500), (0, 300), (800, 333)]), (2, 3, 130, [(0, 64), (64, 200)]),
                                           (1, 12, 2592, [(2592, 2592), (0, 2592)])])
def test_attention_partial_ranges_and_merge(B, H, Lq, ranges):
    """One attention as several launches over disjoint K/V ranges (own keys first, gathered keys later) + m324_attention_merge:
    equals the single launch over all keys, log-sum-exp included."""
    import math
    Lk = sum(n for _, n in ranges)
    g = _gen(Lq + Lk)
    q = torch.randn(B, Lq, H, 64, generator=g).to(DEV).half()
    k = (torch.randn(B, Lk, H, 64, generator=g) * torch.linspace(0.5, 2.0, Lk).view(1, Lk, 1, 1)).to(DEV).half()
    v = torch.randn(B, Lk, H, 64, generator=g).to(DEV).half()
    out = torch.zeros(B * Lq, H * 64, device=DEV, dtype=torch.float16)
    lse = torch.zeros(B * Lq, H, device=DEV)
    parts = len(ranges)
    ws = torch.empty(ops.attention_partial_bytes(B, H, Lq, parts), dtype=torch.uint8, device=DEV)
    kf, vf = k.reshape(B * Lk, H * 64), v.reshape(B * Lk, H * 64)
    for idx, (r0, n) in enumerate(ranges):
        ops.attention(q, kf[r0:], vf[r0:], out, B=B, H=H, Lq=Lq, Lk=n, q_ld=H * 64, k_ld=H * 64, v_ld=H * 64, o_ld=H * 64, q_rows=B * Lq,
                      kv_rows=B * Lk - r0, q_batch_rows=Lq, kv_batch_rows=Lk, scale=0.125, workspace=ws, partial=(parts, idx))
    ops.attention_merge(out, B=B, H=H, Lq=Lq, o_ld=H * 64, parts=parts, workspace=ws, lse=lse, lse_ld=H)
    ref = _attn_ref(q, k, v, 0.125).reshape(B * Lq, H * 64)
    assert torch.isfinite(out).all() and _rel(out, ref) < 1.5e-3, _rel(out, ref)
    s = (q.double().transpose(1, 2) @ k.double().transpose(1, 2).transpose(-2, -1)) * 0.125
    lse_ref = (torch.logsumexp(s, dim=-1) / math.log(2.0)).transpose(1, 2).reshape(B * Lq, H)
    assert float((lse.double() - lse_ref).abs().max()) < 2e-3
    with pytest.raises(Exception):       # a partial launch without a large enough workspace is refused
        ops.attention(q, kf, vf, out, B=B, H=H, Lq=Lq, Lk=Lk, q_ld=H * 64, k_ld=H * 64, v_ld=H * 64, o_ld=H * 64, q_rows=B * Lq,
                      kv_rows=B * Lk, q_batch_rows=Lq, kv_batch_rows=Lk, scale=0.125, workspace=ws[:1024], partial=(parts, 0))


def test_attention_packed_qkv_and_large_scores(attn_mode):
    # packed [rows, 2304] layout (transformer.py:200-202) + large |s| to exercise the lazy-rescale path
    B, H, L = 1, 12, 640
    g = _gen(11)
    qkv = torch.randn(B * L, 3 * H * 64, generator=g)
    qkv[:, :768] *= 3.0
    qkv[300:, 768:1536] *= 4.0   # later keys much larger -> running max moves by > 2^8 after the first tile
    qkv = qkv.to(DEV).half()
    out = torch.zeros(B * L, H * 64, device=DEV, dtype=torch.float16)
    ops.attention(qkv, qkv[:, 768:], qkv[:, 1536:], out, B=B, H=H, Lq=L, Lk=L, q_ld=2304, k_ld=2304, v_ld=2304, o_ld=768,
                  q_rows=B * L, kv_rows=B * L, q_batch_rows=L, kv_batch_rows=L, scale=0.125)
    q, k, v = (qkv[:, i * 768:(i + 1) * 768].reshape(B, L, H, 64) for i in range(3))
    ref = _attn_ref(q, k, v, 0.125).reshape(B * L, H * 64)
    assert _rel(out, ref) < 2e-3, _rel(out, ref)


@pytest.mark.parametrize("T,Lk", [(5, 64), (19, 64), (8, 100), (33, 128)])
def test_attention_shared_query_batches(attn_mode, T, Lk):
    """Decoder pattern: the same queries for every frame (q_batch_rows = 0), <= 128 keys per frame.  With >= 8 frames the kernel
    walks 8 frames per CTA (frame loop: Q resident, one complete attention + epilogue per K/V tile); ragged last group (19 = 8 +
    8 + 3), ragged query tile (300 rows), log-sum-exp per frame."""
    import math
    H, N = 12, 300
    g = _gen(12 + T + Lk)
    q = torch.randn(1, N, H, 64, generator=g).to(DEV).half()
    k = torch.randn(T, Lk, H, 64, generator=g).to(DEV).half()
    v = torch.randn(T, Lk, H, 64, generator=g).to(DEV).half()
    out = torch.zeros(T * N, H * 64, device=DEV, dtype=torch.float16)
    lse = torch.zeros(T * N, H, device=DEV)
    ops.attention(q, k, v, out, B=T, H=H, Lq=N, Lk=Lk, q_ld=768, k_ld=768, v_ld=768, o_ld=768, q_rows=N, kv_rows=T * Lk,
                  q_batch_rows=0, kv_batch_rows=Lk, scale=0.125, lse=lse, lse_ld=H)
    qe = q.expand(T, -1, -1, -1)
    ref = _attn_ref(qe, k, v, 0.125).reshape(T * N, H * 64)
    assert torch.isfinite(out).all() and _rel(out, ref) < 1.5e-3, _rel(out, ref)
    sc = (qe.double().transpose(1, 2) @ k.double().transpose(1, 2).transpose(-2, -1)) * 0.125
    lse_ref = (torch.logsumexp(sc, dim=-1) / math.log(2.0)).transpose(1, 2).reshape(T * N, H)
    assert float((lse.double() - lse_ref).abs().max()) < 2e-3


def test_layernorm_variants():
    rows, C = 1000, 768
    g = _gen(20)
    x = (torch.randn(rows, C, generator=g) * 3 + 1).to(DEV)
    w = (1 + 0.1 * torch.randn(C, generator=g)).to(DEV)
    b = (0.1 * torch.randn(C, generator=g)).to(DEV)
    o16 = torch.empty(rows, 2 * C, device=DEV, dtype=torch.float16)
    o32 = torch.empty(rows, C, device=DEV)
    ops.layernorm(x, w, b, 1e-6, rows, C, out16=o16, ldo16=2 * C, lo_off=C, out32=o32, ldo32=C)
    ref = torch.nn.functional.layer_norm(x.double(), (C,), w.double(), b.double(), 1e-6)
    assert _rel(o32, ref) < 1e-6
    assert _rel(o16[:, :C].float() + o16[:, C:].float(), ref) < 2e-6
    ops.layernorm(x, w, None, 1e-5, rows, C, out32=o32, ldo32=C)
    assert _rel(o32, torch.nn.functional.layer_norm(x.double(), (C,), w.double(), None, 1e-5)) < 1e-6
    # gathered source rows: groups of 64 rows at offset 4 of 324-row frames (Pcd_motion.py:520), bit-exact selection
    xs = torch.randn(3 * 324, C, generator=g).to(DEV)
    o = torch.empty(3 * 64, C, device=DEV)
    ops.layernorm(xs, w, None, 1e-5, 3 * 64, C, src_rpg=64, src_gstride=324, src_goff=4, out32=o, ldo32=C)
    sel = xs.reshape(3, 324, C)[:, 4:68].reshape(-1, C)
    direct = torch.empty(3 * 64, C, device=DEV)
    ops.layernorm(sel.contiguous(), w, None, 1e-5, 3 * 64, C, out32=direct, ldo32=C)
    assert torch.equal(o, direct)


def test_point_features_match_oracle():
    n = 1000
    g = _gen(30)
    xyz = (torch.rand(n, 3, generator=g) - 0.5)
    nrm = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
    rgb = torch.rand(n, 3, generator=g)
    a0 = torch.empty(n, 128, device=DEV, dtype=torch.float16)
    ops.point_embed_features(xyz.to(DEV), n, a0, 128, 64)
    proj = torch.einsum("nd,de->ne", xyz.double(), orc.point_embed_basis().double())
    emb = torch.cat([proj.sin(), proj.cos(), xyz.double()], dim=1)
    got = a0[:, :64].double() + a0[:, 64:].double()
    assert (got[:, 51:] == 0).all()
    # sin/cos of fp32 arguments up to ~200 rad: absolute error of the fp32 product x*2^k*pi dominates (~1e-5)
    assert float((got[:, :51].cpu() - emb).abs().max()) < 5e-5
    a1 = torch.zeros(n, 1664, device=DEV, dtype=torch.float16)
    ops.point_extra_features(nrm.to(DEV), rgb.to(DEV), n, a1, 1664, 768, 832, 832)
    got = (a1[:, 768:832].double() + a1[:, 1600:1664].double()).cpu()
    assert float((got[:, :3] - nrm.double()).abs().max()) < 1e-6 and float((got[:, 3:6] - rgb.double()).abs().max()) < 1e-6
    assert (got[:, 6:] == 0).all()


@pytest.mark.parametrize("Hin,Win", [(224, 224), (160, 192), (720, 720)])
def test_preprocess_matches_interpolate_normalise_im2col(Hin, Win):
    F_ = 2
    g = _gen(Hin)
    video = torch.rand(F_, Hin, Win, 3, generator=g).to(DEV)
    patches = torch.empty(F_ * 256, 640, device=DEV, dtype=torch.float16)
    ops.preprocess_frames(video, F_, Hin, Win, 224, patches, 640, 640)
    x = torch.nn.functional.interpolate(video.permute(0, 3, 1, 2).double(), (224, 224), mode="bilinear", align_corners=False)
    mean = torch.tensor([0.485, 0.456, 0.406], device=DEV, dtype=torch.float64).view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225], device=DEV, dtype=torch.float64).view(1, 3, 1, 1)
    x = (x - mean) / std
    ref = x.reshape(F_, 3, 16, 14, 16, 14).permute(0, 2, 4, 1, 3, 5).reshape(F_ * 256, 588)
    assert (patches[:, 588:] == 0).all()
    assert float((patches[:, :588].double() - ref).abs().max()) < 2e-3  # fp16 storage of values up to ~2.6
    if Hin == 224:  # identity resize: same fp32 arithmetic as the reference's (x - mean) / std, then the fp16 store
        x32 = (video.permute(0, 3, 1, 2) - mean.float()) / std.float()
        ref32 = x32.reshape(F_, 3, 16, 14, 16, 14).permute(0, 2, 4, 1, 3, 5).reshape(F_ * 256, 588)
        assert torch.equal(patches[:, :588], ref32.half())


def test_dino_assemble_and_token_assembly():
    F_, C = 3, 768
    g = _gen(40)
    patch = torch.randn(F_ * 256, C, generator=g).to(DEV)
    cls = torch.randn(C, generator=g).to(DEV)
    pos = torch.randn(257, C, generator=g).to(DEV)
    x = torch.empty(F_ * 257, C, device=DEV)
    ops.dino_assemble(patch, cls, pos, F_, 256, C, x)
    ref = torch.cat([cls.view(1, 1, C).expand(F_, 1, C), patch.view(F_, 256, C)], dim=1) + pos.view(1, 257, C)
    assert torch.equal(x.view(F_, 257, C), ref)
    B, T = 1, 3
    nw = (1 + 0.1 * torch.randn(C, generator=g)).to(DEV)
    nb = (0.1 * torch.randn(C, generator=g)).to(DEV)
    pe = torch.randn(T * 256, C, generator=g).to(DEV)
    sp0, spr = torch.randn(4, C, generator=g).to(DEV), torch.randn(4, C, generator=g).to(DEV)
    mesh = torch.randn(B * 64, C, generator=g).to(DEV)
    lw = (1 + 0.1 * torch.randn(C, generator=g)).to(DEV)
    out = torch.empty(B * T * 324, C, device=DEV)
    ops.assemble_tokens(x, nw, nb, 1e-6, pe, sp0, spr, mesh, lw, 1e-5, B, T, 64, 256, C, out)
    xd = x.double().view(T, 257, C)
    vid = torch.nn.functional.layer_norm(xd, (C,), nw.double(), nb.double(), 1e-6)[:, 1:] + pe.double().view(T, 256, C)
    spec = torch.stack([sp0.double()] + [spr.double()] * (T - 1))
    tok = torch.cat([spec, mesh.double().view(1, 64, C).expand(T, 64, C), vid], dim=1)
    ref = torch.nn.functional.layer_norm(tok, (C,), lw.double(), None, 1e-5).reshape(T * 324, C)
    assert _rel(out, ref) < 1e-6


def test_head3_and_mse():
    rows, C = 5000, 768
    g = _gen(50)
    h = torch.randn(rows, C, generator=g).to(DEV)
    w3 = (0.02 * torch.randn(3, C, generator=g)).to(DEV)
    b3 = (0.1 * torch.randn(3, generator=g)).to(DEV)
    tgt = torch.randn(rows, 3, generator=g).to(DEV)
    out = torch.empty(rows, 3, device=DEV)
    partials = torch.zeros(4096, device=DEV)
    loss = torch.zeros(2, device=DEV)
    n = ops.head3_mse(h, C, w3, b3, rows, C, out, tgt, partials)
    ops.mse_finalize(partials, n, rows * 3, 0.5, loss)
    ref = h.double() @ w3.double().t() + b3.double()
    assert _rel(out, ref) < 1e-6
    mse = ((ref - tgt.double()) ** 2).mean()
    assert abs(float(loss[0]) - float(mse)) < 1e-6 * float(mse) + 1e-9
    assert abs(float(loss[1]) - 0.5 * float(mse)) < 1e-6 * float(mse) + 1e-9
    # standalone loss kernel == model/loss.py:59-61
    loss2 = torch.zeros(2, device=DEV)
    ops.mse_loss(out, tgt, rows * 3, 1.0, partials, loss2)
    assert abs(float(loss2[0]) - float(torch.nn.functional.mse_loss(out, tgt))) < 1e-6
    # bit-reproducible (fixed grid, ordered final reduce)
    loss3 = torch.zeros(2, device=DEV)
    ops.mse_loss(out, tgt, rows * 3, 1.0, partials, loss3)
    assert torch.equal(loss2, loss3)


@pytest.mark.parametrize("M", [4096, 333])
def test_head_fused_into_the_gemm_epilogue(M):
    """shared_mlp_output.3 folded into the epilogue of shared_mlp_output.1 (gemm(head_w=, head_part=) + head3_from_partials): the
    3-pass split GEMM + bias + GELU never writes its [M, 768] result; out / MSE equal the unfused sequence (GEMM -> fp32 hidden ->
    head3_mse) up to the fp32 summation order, and the fp64 reference."""
    N = K = 768
    g = _gen(60 + M)
    A = torch.randn(M, K, generator=g).to(DEV)
    W = (0.05 * torch.randn(N, K, generator=g)).to(DEV)
    b1 = (0.1 * torch.randn(N, generator=g)).to(DEV)
    w3 = (0.02 * torch.randn(3, N, generator=g)).to(DEV)
    b3 = (0.1 * torch.randn(3, generator=g)).to(DEV)
    tgt = torch.randn(M, 3, generator=g).to(DEV)
    A16 = torch.empty(M, 2 * K, device=DEV, dtype=torch.float16)
    W16 = torch.empty(N, 2 * K, device=DEV, dtype=torch.float16)
    ops.cast_pad_f16(A, M, K, A16, 2 * K, K, lo_off=K)
    ops.cast_pad_f16(W, N, K, W16, 2 * K, K, lo_off=K)
    kw = dict(passes=3, a_lo_off=K, w_lo_off=K, bias=b1, act=1)
    # unfused
    hbuf = torch.empty(M, N, device=DEV)
    out0, part0, loss0 = torch.empty(M, 3, device=DEV), torch.zeros(4096, device=DEV), torch.zeros(2, device=DEV)
    ops.gemm(A16, W16, M, N, K, out32=hbuf, ldo32=N, **kw)
    n0 = ops.head3_mse(hbuf, N, w3, b3, M, N, out0, tgt, part0)
    ops.mse_finalize(part0, n0, M * 3, 1.0, loss0)
    # fused
    hpart = torch.full((M, N // 64, 4), float("nan"), device=DEV)
    out1, part1, loss1 = torch.empty(M, 3, device=DEV), torch.zeros(4096, device=DEV), torch.zeros(2, device=DEV)
    ops.gemm(A16, W16, M, N, K, head_w=w3, head_part=hpart, **kw)
    n1 = ops.head3_from_partials(hpart, N // 64, b3, M, out1, tgt, part1)
    ops.mse_finalize(part1, n1, M * 3, 1.0, loss1)
    torch.cuda.synchronize()
    assert torch.isfinite(hpart[:, :, :3]).all()
    ref = torch.nn.functional.gelu(A.double() @ W.double().t() + b1.double()) @ w3.double().t() + b3.double()
    assert _rel(out1, ref) < 2e-5 and _rel(out0, ref) < 2e-5, (_rel(out1, ref), _rel(out0, ref))
    assert _rel(out1, out0) < 2e-6
    assert abs(float(loss1[0]) - float(loss0[0])) < 1e-5 * float(loss0[0])
    with pytest.raises(Exception):
        ops.gemm(A16, W16, M, N, K, head_w=w3, **kw)          # head_w without head_part


@pytest.mark.parametrize("M,N,K,mod,div", [(4096 * 3, 768, 768, 4096, 0), (1000, 768, 128, 320, 0), (2048, 256, 64, 64, 512), (700, 768, 192, 0, 0)])
def test_gemm_modulo_residual_prefetched_by_tma(M, N, K, mod, div):
    """The decoder's per-point residual (the same feature row for every frame, Pcd_motion.py:556-560): out = A W^T + resid[map(row)], map(row) =
    (row / div) * mod + row % mod.  The 2-CTA kernel prefetches the residual as 32 x 32 TMA boxes when boxes map to consecutive rows; result
    bit-identical to the transposed-LDG path (knob 5 = 1), ragged last row tile included, and equal to the fp64 reference."""
    g = _gen(M + N + K + mod)
    A = torch.randn(M, K, generator=g).to(DEV).half()
    W = (0.05 * torch.randn(N, K, generator=g)).to(DEV).half()
    rrows = ((M - 1) // div + 1) * mod if (mod and div) else (mod if mod else M)
    resid = torch.randn(rrows, N, generator=g).to(DEV)
    out_f, out_s = torch.empty(M, N, device=DEV), torch.empty(M, N, device=DEV)
    kw = dict(resid=resid, ldr=N, resid_mod=mod, resid_div=div, ldo32=N)
    ops.gemm(A, W, M, N, K, out32=out_f, **kw)
    ops.set_tuning(5, 1)
    try:
        ops.gemm(A, W, M, N, K, out32=out_s, **kw)
    finally:
        ops.set_tuning(5, 0)
    torch.cuda.synchronize()
    rows = torch.arange(M, device=DEV)
    idx = ((rows // div) * mod + rows % mod) if (mod and div) else (rows % mod if mod else rows)
    ref = A.double() @ W.double().t() + resid.double()[idx]
    assert _rel(out_f, ref) < 1e-5 and _rel(out_s, ref) < 1e-5
    assert torch.equal(out_f, out_s)
