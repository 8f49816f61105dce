"""Decoder cross-attention launch timed alone: T frames x N shared queries x 64 keys (frame loop on / off via knob 0 = 1)."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from motion324_b200 import ops
dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
T, N, H, M, d = 32, 4096, 12, 64, 768
q = torch.randn(N, d, device=dev).half()
kv = torch.randn(T * M, 2 * d, device=dev).half()
o = torch.empty(T * N, d, device=dev, dtype=torch.float16)
def run():
    ops.attention(q, kv, kv[:, d:], o, B=T, H=H, Lq=N, Lk=M, q_ld=d, k_ld=2 * d, v_ld=2 * d, o_ld=d, q_rows=N, kv_rows=T * M,
                  q_batch_rows=0, kv_batch_rows=M, scale=0.125)
for knob in (0, 1):
    ops.set_tuning(0, knob)
    for _ in range(3): run()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(10):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); run(); e.record(); e.synchronize(); tot += s.elapsed_time(e)
    print(json.dumps(dict(frame_loop=knob == 0, ms=tot / 10, out_gbs=T * N * d * 2 / (tot / 10 * 1e-3) / 1e9)))
