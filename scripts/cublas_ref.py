"""cuBLAS (torch.matmul, fp16) at the trunk GEMM shapes, timed like scripts/gemm_sweep.py: the library baseline."""
import torch
dev = "cuda"
L, d = 32 * 324, 768
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, cold, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); tot = 0.0
    for _ in range(iters):
        if cold: flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); e.synchronize(); tot += s.elapsed_time(e)
    return tot / iters * 1e3
for name, M, N, K in [("qkv", L, 3 * d, d), ("up", L, 4 * d, d), ("down", L, d, 4 * d), ("fc", L, d, d), ("big", 8192, 8192, 8192)]:
    A = torch.randn(M, K, device=dev).half(); W = torch.randn(N, K, device=dev).half(); out = torch.empty(M, N, device=dev, dtype=torch.float16)
    fn = lambda: torch.matmul(A, W.t(), out=out)
    w, c = timeit(fn, False), timeit(fn, True)
    fl = 2.0 * M * N * K
    print(f"cublas {name:5s} M={M} N={N} K={K}: warm {w:8.1f} us {fl / w / 1e6:7.0f} TF/s   cold {c:8.1f} us {fl / c / 1e6:7.0f} TF/s")
