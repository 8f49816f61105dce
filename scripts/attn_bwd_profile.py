"""Attention backward launches only (ncu target):  ncu --set full --clock-control none --import-source on -k regex:attn_bwd -s 1 -c 1 ..."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from motion324_b200 import ops
d, H = 768, 12
B, L = int(os.environ.get("M324_B", "4")), int(os.environ.get("M324_L", "3888"))
rows = B * L
qkv = torch.randn(rows, 3 * d, device="cuda").half()
o = torch.empty(rows, d, device="cuda", dtype=torch.float16)
lse = torch.empty(rows, H, device="cuda")
kw = dict(B=B, H=H, Lq=L, Lk=L, q_ld=3 * d, k_ld=3 * d, v_ld=3 * d, q_rows=rows, kv_rows=rows, q_batch_rows=L, kv_batch_rows=L, scale=0.125)
ops.attention(qkv, qkv[:, d:], qkv[:, 2 * d:], o, o_ld=d, lse=lse, lse_ld=H, **kw)
dO = (torch.randn(rows, d, device="cuda") * 0.01).half()
D = torch.empty(rows, H, device="cuda")
ops.attn_dot(dO, d, o, d, rows, H, D, H)
dqkv = torch.zeros(rows, 3 * d, device="cuda")
for _ in range(3):
    ops.attention_bwd(qkv, qkv[:, d:], qkv[:, 2 * d:], dO, lse, D, dqkv, dqkv[:, d:], dqkv[:, 2 * d:], do_ld=d, lse_ld=H, d_ld=H,
                      dq_ld=3 * d, dk_ld=3 * d, dv_ld=3 * d, **kw)
torch.cuda.synchronize()
print("ok")
