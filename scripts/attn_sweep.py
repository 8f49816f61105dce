"""Attention kernel timings at the model's shapes (CUDA events, L2 flushed)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from motion324_b200 import ops  # noqa: E402

dev = "cuda"
d, H = 768, 12
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); e.synchronize()
        tot += s.elapsed_time(e)
    return tot / iters * 1e3


def self_attn(B, L):
    qkv = torch.randn(B * L, 3 * d, device=dev).half()
    o = torch.empty(B * L, d, device=dev, dtype=torch.float16)
    return lambda: ops.attention(qkv, qkv[:, d:], qkv[:, 2 * d:], o, B=B, H=H, Lq=L, Lk=L, q_ld=3 * d, k_ld=3 * d, v_ld=3 * d,
                                 o_ld=d, q_rows=B * L, kv_rows=B * L, q_batch_rows=L, kv_batch_rows=L, scale=0.125)


def dec_attn(T, N, M):
    q = torch.randn(N, d, device=dev).half()
    kv = torch.randn(T * M, 2 * d, device=dev).half()
    o = torch.empty(T * N, d, device=dev, dtype=torch.float16)
    return lambda: ops.attention(q, kv, kv[:, d:], o, B=T, H=H, Lq=N, Lk=M, q_ld=d, k_ld=2 * d, v_ld=2 * d, o_ld=d, q_rows=N,
                                 kv_rows=T * M, q_batch_rows=0, kv_batch_rows=M, scale=0.125)


import itertools
for ev, skew in [(0, 0), (0, 1)]:
  ops.set_tuning(0, ev); ops.set_tuning(1, skew)
  print(f"--- attention mode {ev} (0/1 = pair kernel, 2 = K/V-split kernel), knob 1 = {skew} (1 = MUFU turn-taking off)")
  for name, fn, fl in [("global 1x10368", self_attn(1, 10368), 4.0 * H * 10368 * 10368 * 64),
                     ("local 32x324", self_attn(32, 324), 4.0 * 32 * H * 324 * 324 * 64),
                     ("dino 32x257", self_attn(32, 257), 4.0 * 32 * H * 257 * 257 * 64),
                     ("latent 1x64", self_attn(1, 64), 4.0 * H * 64 * 64 * 64),
                     ("decoder 32x(4096x64)", dec_attn(32, 4096, 64), 4.0 * 32 * H * 4096 * 64 * 64),
                     ("global T=128 1x41472", self_attn(1, 41472), 4.0 * H * 41472 * 41472 * 64)]:
      us = timeit(fn)
      print(f"{name:24s} {us:10.1f} us {fl / us / 1e6:8.1f} TFLOP/s")
