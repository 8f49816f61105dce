"""Attention backward kernel timed alone (CUDA events, L2 flushed): global / local trunk shapes of the training step."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from motion324_b200 import ops

dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
d, H = 768, 12


def t(fn, iters=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); e.synchronize()
        tot += s.elapsed_time(e)
    return tot / iters


TUNES = [int(x) for x in os.environ.get('M324_BWD_TUNES', '0').split(',')]
for tune, (B, L) in [(tu, bl) for tu in TUNES for bl in ((32, 3888), (384, 324), (1, 10368))]:
    ops.set_tuning(3, tune)
    rows = B * L
    qkv = torch.randn(rows, 3 * d, device=dev).half()
    o = torch.empty(rows, d, device=dev, dtype=torch.float16)
    lse = torch.empty(rows, H, device=dev)
    kw = dict(B=B, H=H, Lq=L, Lk=L, q_ld=3 * d, k_ld=3 * d, v_ld=3 * d, q_rows=rows, kv_rows=rows, q_batch_rows=L, kv_batch_rows=L, scale=0.125)
    ops.attention(qkv, qkv[:, d:], qkv[:, 2 * d:], o, o_ld=d, lse=lse, lse_ld=H, **kw)
    dO = (torch.randn(rows, d, device=dev) * 0.01).half()
    D = torch.empty(rows, H, device=dev)
    ops.attn_dot(dO, d, o, d, rows, H, D, H)
    dqkv = torch.zeros(rows, 3 * d, device=dev)
    ms = t(lambda: ops.attention_bwd(qkv, qkv[:, d:], qkv[:, 2 * d:], dO, lse, D, dqkv, dqkv[:, d:], dqkv[:, 2 * d:], do_ld=d, lse_ld=H, d_ld=H,
                                     dq_ld=3 * d, dk_ld=3 * d, dv_ld=3 * d, **kw))
    fl = 10.0 * B * H * L * L * 64
    print(json.dumps(dict(tune=tune, B=B, L=L, ms=ms, tflops=fl / ms / 1e9)))
