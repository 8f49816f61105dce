"""A/B of the attention tail split (work items of the last, partly filled wave cut into K/V ranges): global-layer shapes."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from motion324_b200 import ops

dev = torch.device("cuda")
if os.environ.get('M324_KNOB1'):
    ops.set_tuning(1, int(os.environ['M324_KNOB1']))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
d, H = 768, 12


def t(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); e.synchronize()
        tot += s.elapsed_time(e)
    return tot / iters


for T in [int(x) for x in os.environ.get('M324_TS', '8,12,16,24,32,48,64,128').split(',')]:
    L = T * 324
    qkv = torch.randn(L, 3 * d, device=dev).half()
    o = torch.empty(L, d, device=dev, dtype=torch.float16)
    kw = dict(B=1, H=H, Lq=L, Lk=L, q_ld=3 * d, k_ld=3 * d, v_ld=3 * d, o_ld=d, q_rows=L, kv_rows=L, q_batch_rows=L, kv_batch_rows=L, scale=0.125)
    a = t(lambda: ops.attention(qkv, qkv[:, d:], qkv[:, 2 * d:], o, workspace=None, **kw))
    b = t(lambda: ops.attention(qkv, qkv[:, d:], qkv[:, 2 * d:], o, **kw))
    fl = 4.0 * H * L * L * 64
    items = ((L + 255) // 256) * H
    print(json.dumps(dict(T=T, L=L, items=items, waves=items / 148, ms_nosplit=a, ms_split=b, tflops_nosplit=fl / a / 1e9, tflops_split=fl / b / 1e9)))
