import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from motion324_b200 import ops
DEV = "cuda"
def rel(a, b):
    a, b = a.double(), b.double(); return float((a - b).norm() / b.norm())
M, N, K = 256, 768, 768
for adt, wdt in [(torch.float16, torch.float16), (torch.bfloat16, torch.bfloat16), (torch.bfloat16, torch.float16)]:
    A = (torch.randn(M, K) ).to(DEV).to(adt); W = (torch.randn(N, K) * 0.05).to(DEV).to(wdt)
    ref = (A.double().cpu() @ W.double().cpu().t())
    out = torch.zeros(M, N, device=DEV)
    try:
        ops.gemm(A, W, M, N, K, out32=out, ldo32=N); torch.cuda.synchronize()
        print(adt, wdt, "rel", rel(out.cpu(), ref))
    except Exception as e:
        print(adt, wdt, "ERR", e); break
