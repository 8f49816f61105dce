"""Global-layer attention launches only (Lq = Lk = T*324, H = 12): the target of an `ncu --set full --import-source on` capture.

    ncu --set full --clock-control none --import-source on -k regex:attn -s 2 -c 1 -o gpurun_out/attn python scripts/attn_profile.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from motion324_b200 import ops  # noqa: E402

T = int(os.environ.get("M324_T", "32"))
mode = int(os.environ.get("M324_ATTN_MODE", "0"))
d, H, L = 768, 12, T * 324
if mode:
    ops.set_tuning(0, mode)
qkv = torch.randn(L, 3 * d, device="cuda").half()
o = torch.empty(L, d, device="cuda", dtype=torch.float16)
for _ in range(4):
    ops.attention(qkv, qkv[:, d:], qkv[:, 2 * d:], o, B=1, H=H, Lq=L, Lk=L, q_ld=3 * d, k_ld=3 * d, v_ld=3 * d, o_ld=d,
                  q_rows=L, kv_rows=L, q_batch_rows=L, kv_batch_rows=L, scale=0.125)
torch.cuda.synchronize()
print("ok", float(o.float().abs().mean()))
