"""GEMM variants at the trunk shapes: separates mainloop from epilogue cost (CUDA events, warm and cold L2)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from motion324_b200 import ops  # noqa: E402

dev = "cuda"
d, L = 768, 32 * 324
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, cold, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        if cold:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); e.synchronize()
        tot += s.elapsed_time(e)
    return tot / iters * 1e3  # us


A768 = torch.randn(L, d, device=dev).half()
A3072 = torch.randn(L, 3072, device=dev).half()
W_qkv = (torch.randn(3 * d, d, device=dev) * 0.02).half()
W_up = (torch.randn(3072, d, device=dev) * 0.02).half()
W_dn = (torch.randn(d, 3072, device=dev) * 0.02).half()
W_fc = (torch.randn(d, d, device=dev) * 0.02).half()
o16 = torch.empty(L, 3072, device=dev, dtype=torch.float16)
x = torch.zeros(L, d, device=dev)
qn = torch.ones(64, device=dev)
cases = {
    "qkv  plain16      ": lambda m: ops.gemm(A768, W_qkv, L, 3 * d, d, out16=o16, ldo16=3072, force_bn128=m),
    "qkv  qknorm16     ": lambda m: ops.gemm(A768, W_qkv, L, 3 * d, d, out16=o16, ldo16=3072, qn_w=qn, kn_w=qn, qk_cols=d, force_bn128=m),
    "up   plain16      ": lambda m: ops.gemm(A768, W_up, L, 3072, d, out16=o16, ldo16=3072, force_bn128=m),
    "up   gelu16       ": lambda m: ops.gemm(A768, W_up, L, 3072, d, act=1, out16=o16, ldo16=3072, force_bn128=m),
    "down plain32      ": lambda m: ops.gemm(A3072, W_dn, L, d, 3072, out32=x, ldo32=d, force_bn128=m),
    "down resid32      ": lambda m: ops.gemm(A3072, W_dn, L, d, 3072, resid=x, ldr=d, out32=x, ldo32=d, force_bn128=m),
    "fc   resid32      ": lambda m: ops.gemm(A768, W_fc, L, d, d, resid=x, ldr=d, out32=x, ldo32=d, force_bn128=m),
}
flops = {"qkv": 2.0 * L * 3 * d * d, "up ": 2.0 * L * 3072 * d, "dow": 2.0 * L * 3072 * d, "fc ": 2.0 * L * d * d}
print(f"{'case':20s} {'mode':>14s} {'warm us':>9s} {'TF/s':>7s} {'cold us':>9s} {'TF/s':>7s}")
for name, fn in cases.items():
    fl = flops[name[:3]]
    for mode, mname in ((2, "1cta"), (4, "2cta"), (16 + 2, "1cta-noepi"), (16 + 4, "2cta-noepi"), (48 + 2, "1cta-noepi-noload"), (48 + 4, "2cta-noepi-noload")):
        w = timeit(lambda: fn(mode), False)
        c = timeit(lambda: fn(mode), True)
        print(f"{name:20s} {mname:>14s} {w:9.1f} {fl / w / 1e6:7.0f} {c:9.1f} {fl / c / 1e6:7.0f}")
