"""Stand-alone launches of the hot kernels at the bench shapes (for `ncu --set full -k regex:...`)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from motion324_b200 import ops  # noqa: E402

dev = "cuda"
d, H, L = 768, 12, 32 * 324
which = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
qkv = torch.randn(L, 3 * d, device=dev).half()
o = torch.empty(L, d, device=dev, dtype=torch.float16)
A = torch.randn(L, 3072, device=dev).half()
W1 = (torch.randn(3072, d, device=dev) * 0.02).half()
W2 = (torch.randn(d, 3072, device=dev) * 0.02).half()
Wq = (torch.randn(3 * d, d, device=dev) * 0.02).half()
hid = torch.empty(L, 3072, device=dev, dtype=torch.float16)
x = torch.zeros(L, d, device=dev)
qn = torch.ones(64, device=dev)
for _ in range(reps):
    if which in ("all", "attn"):
        ops.attention(qkv, qkv[:, d:], qkv[:, 2 * d:], o, B=1, H=H, Lq=L, Lk=L, q_ld=3 * d, k_ld=3 * d, v_ld=3 * d, o_ld=d,
                      q_rows=L, kv_rows=L, q_batch_rows=L, kv_batch_rows=L, scale=0.125)
    if which in ("all", "gemm"):
        ops.gemm(A, W1, L, 3072, d, lda=3072, act=1, out16=hid, ldo16=3072)          # MLP up + GELU
        ops.gemm(A, W2, L, d, 3072, resid=x, ldr=d, out32=x, ldo32=d)                # MLP down + residual
        ops.gemm(A, Wq, L, 3 * d, d, lda=3072, out16=hid, ldo16=3072, qn_w=qn, kn_w=qn, qk_cols=d)  # to_qkv + qk-norm
torch.cuda.synchronize()
print("done")
