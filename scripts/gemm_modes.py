"""GEMM tile-shape modes at given M: time per mode (1 = 1-CTA 128x128, 2 = 1-CTA 128x256, 3 = 2-CTA 256x128, 4 = 2-CTA 256x256, 0 = auto)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from motion324_b200 import ops
dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); tot = 0.0
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); e.synchronize(); tot += s.elapsed_time(e)
    return tot / iters * 1e3
for M in (8224, 10368):
    for (N, K, kind) in ((2304, 768, "qkv16"), (768, 768, "proj-resid32"), (3072, 768, "up-gelu16"), (768, 3072, "down-resid32")):
        A = torch.randn(M, K, device=dev).half(); W = (torch.randn(N, K, device=dev) * 0.02).half()
        o16 = torch.empty(M, N, device=dev, dtype=torch.float16); x = torch.zeros(M, N, device=dev); b = torch.zeros(N, device=dev)
        line = f"M={M} N={N} K={K} {kind:14s}"
        for mode in (0, 1, 2, 3, 4):
            if kind.endswith("16"):
                fn = lambda: ops.gemm(A, W, M, N, K, bias=b, act=1 if "gelu" in kind else 0, out16=o16, ldo16=N, force_bn128=mode)
            else:
                fn = lambda: ops.gemm(A, W, M, N, K, bias=b, resid=x, ldr=N, out32=x, ldo32=N, force_bn128=mode)
            line += f"  m{mode}:{timeit(fn):6.1f}us"
        print(line)
