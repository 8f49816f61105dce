"""Training-step timing: forward with saved activations + hand-written backward of
Motion_Latent_Model on libm324, optionally + the single gradient all-reduce and a fused AdamW step (train.py:157-213).

    python scripts/train_bench.py [--batch 32] [--frames 12] [--points 4096] [--steps 5] [--optimizer] [--profile]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/train_bench.py ...   (data parallel)

Prints one JSON line on rank 0: ms/step (CUDA events, max over ranks), frames/s (whole job), effective TFLOP/s counted as
forward FLOPs (SURVEY.md 8d formulas) + 2 x the trainable part (frozen DINOv2 has no backward), all-reduce bytes and time.
--profile brackets ONE step with cudaProfilerStart/Stop (ncu --profile-from-start off).
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from motion324_b200.model.Pcd_motion import Motion_Latent_Model  # noqa: E402
from motion324_b200.utils.config import make_config  # noqa: E402
from motion324_b200 import ops  # noqa: E402
from motion324_b200.utils import synthetic as orc  # seeded weights / inputs generator


def flops(B, T, N, S, d=768, L=324):
    f_shape = 2 * (S * (72 + 51 * d + 774 * d) + 2 * S * d * d + 128 * d * d + 128 * S * d + 512 * d * d + 4 * (64 * 12 * d * d + 2 * 64 * 64 * d))
    f_dino = T * 2 * (256 * 588 * d + 12 * (257 * 12 * d * d + 2 * 257 * 257 * d))
    f_trunk = 16 * T * L * 12 * d * d * 2 + 8 * T * 4 * L * L * d + 8 * 4 * (T * L) ** 2 * d
    f_dec = T * 2 * (N * (2 * d * d + 8 * d * d + d * d + 3 * d) + 128 * d * d + 128 * N * d) + N * (72 + 825 * d) * 2
    fwd = B * (f_shape + f_dino + f_trunk + f_dec)
    return fwd, fwd + 2 * B * (f_shape + f_trunk + f_dec)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--frames", type=int, default=12)
    ap.add_argument("--points", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--optimizer", action="store_true")
    ap.add_argument("--profile", action="store_true")
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=dev)
    B, T, N = a.batch, a.frames, a.points
    model = Motion_Latent_Model(make_config(frames=T, drop_rate=0.1))
    model.load_state_dict(orc.init_state_dict(0, dict(frames=T)), strict=True)
    model = model.to(dev)
    model.train()
    one = orc.make_inputs(seed=1 + rank, B=1, T=T, N=N, S=N)
    sample = {k: v.to(dev).expand(B, *v.shape[1:]).contiguous() for k, v in one.items()}
    sample["point_clouds"] = sample["point_clouds"] + 0.01 * torch.randn_like(sample["point_clouds"])
    opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-5, fused=True) if a.optimizer else None
    gb = model.grad_buffer()
    ar_ms = []

    def step():
        ret = model.forward_backward(sample)
        if world > 1:      # the single gradient all-reduce of train.py's DDP (C1), on the flat buffer
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            torch.distributed.all_reduce(gb.flat)
            gb.flat.mul_(1.0 / world)
            e.record()
            ar_ms.append((s, e))
        if opt is not None:
            opt.step()
        return ret

    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    ar_ms.clear()
    ops.LAUNCHES[0] = 0
    if a.profile:
        torch.cuda.profiler.start()
        ret = step()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        print("loss", float(ret.loss_metrics.loss), "launches", ops.LAUNCHES[0])
        return
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(a.steps):
        ret = step()
    e.record()
    torch.cuda.synchronize()
    ms = torch.tensor([s.elapsed_time(e) / a.steps], device=dev)
    if world > 1:
        torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
    fwd, tot = flops(B, T, N, N)
    if rank == 0:
        ms = float(ms[0])
        print(json.dumps({
            "what": "training step (forward + backward" + (" + grad all-reduce" if world > 1 else "") + (" + fused AdamW" if opt else "") + ")",
            "batch_per_gpu": B, "frames": T, "points": N, "n_gpus": world, "ms_per_step": ms,
            "frames_per_s": world * B * T / (ms * 1e-3), "tflops_effective_per_gpu": tot / (ms * 1e-3) / 1e12,
            "fwd_tflop": fwd / 1e12, "fwd_bwd_tflop": tot / 1e12, "launches_per_step": ops.LAUNCHES[0] / a.steps,
            "allreduce_bytes": gb.flat.numel() * 4 if world > 1 else 0,
            "allreduce_ms": (sum(x.elapsed_time(y) for x, y in ar_ms) / len(ar_ms)) if ar_ms else None,
            "loss": float(ret.loss_metrics.loss), "mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
