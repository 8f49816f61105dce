"""BASELINE.json configs[4]: long-sequence sweep, frames T x points N, forward + loss of ONE clip, on 1 / 2 / 4 / 8 GPUs.

    python scripts/sweep.py                                                      # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/sweep.py     # frames sharded over N GPUs

One JSON line per grid point on rank 0 (CUDA events, max over ranks): ms per clip, frames/s, the algorithmic FLOPs of the clip
(SURVEY.md 8(d) formulas; the reference's T-fold recomputation of the decoder point embedding is NOT counted), the effective
TFLOP/s per GPU and its fraction of the measured sustained bf16 GEMM peak (MEASURED_PEAKS.json; the kernels are timed inside a
long step), and the share of the FLOPs that is global attention (the part that is quadratic in T).  On N > 1 GPUs the clip's
frames are sharded (Motion_Latent_Model.frame_parallel: strong scaling, one K|V all-gather per global layer); grid points whose
frame count does not divide by the world size are skipped.
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
from motion324_b200.utils import synthetic as syn  # noqa: E402


def flops(T, N, S, d=768, L=324):
    f_shape = 2 * (S * (72 + 51 * d + 774 * d) + 2 * S * d * d + 128 * d * d + 128 * S * d + 512 * d * d + 4 * (64 * 12 * d * d + 2 * 64 * 64 * d))
    f_dino = T * 2 * (256 * 588 * d + 12 * (257 * 12 * d * d + 2 * 257 * 257 * d))
    f_global = 8 * 4 * (T * L) ** 2 * d
    f_trunk = 16 * T * L * 12 * d * d * 2 + 8 * T * 4 * L * L * d + f_global
    f_dec = T * 2 * (N * (2 * d * d + 8 * d * d + d * d + 3 * d) + 128 * d * d + 128 * N * d) + N * (72 + 825 * d) * 2
    return f_shape + f_dino + f_trunk + f_dec, f_global


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, nargs="+", default=[8, 32, 128, 512])
    ap.add_argument("--points", type=int, nargs="+", default=[1024, 4096, 16384])
    ap.add_argument("--shape-samples", type=int, default=4096)
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=dev)
    peak = 1400.7
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak = json.load(f).get("bf16_tflops_sustained", peak)
    except Exception:
        pass
    S = a.shape_samples
    for T in a.frames:
        if T % world != 0:
            continue
        model = Motion_Latent_Model(make_config(frames=T))
        model.load_state_dict(syn.init_state_dict(0, dict(frames=T)), strict=True)
        model = model.to(dev)
        model.eval()
        model.frame_parallel(world > 1)
        g = torch.Generator().manual_seed(2)
        video = torch.rand(1, T, 224, 224, 3, generator=g).to(dev)
        for N in a.points:
            one = syn.make_inputs(seed=1, B=1, T=1, N=N, S=S)
            sample = {k: v.to(dev) for k, v in one.items()}
            sample["rgb_video"] = video
            sample["point_clouds"] = sample["ref_pcd"][:, None] + 0.05 * torch.randn(1, T, N, 3, generator=g).to(dev)
            for _ in range(2):
                model(sample)
            if world > 1:
                torch.distributed.barrier()
            torch.cuda.synchronize()
            reps = 2 if T >= 512 else 3 if T >= 128 else 6
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(reps):
                ret = model(sample)
            e.record()
            torch.cuda.synchronize()
            ms = torch.tensor([s.elapsed_time(e) / reps], device=dev)
            if world > 1:
                torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
            ms = float(ms[0])
            fl, fl_glob = flops(T, N, S)
            if rank == 0:
                tf = fl / (ms * 1e-3) / 1e12
                print(json.dumps(dict(T=T, N=N, S=S, n_gpus=world, mode="frames sharded (strong scaling)" if world > 1 else "one GPU",
                                      ms_per_clip=ms, frames_per_s=T / ms * 1e3, tflop_per_clip=fl / 1e12, global_attention_flop_share=fl_glob / fl,
                                      tflops_effective_total=tf, tflops_effective_per_gpu=tf / world, peak_tflops_sustained_measured=peak,
                                      roofline_frac_per_gpu=tf / world / peak, loss=float(ret.loss_metrics.loss),
                                      peak_gib=torch.cuda.max_memory_allocated() / 2 ** 30)), flush=True)
            del sample, ret
        del model, video
        torch.cuda.empty_cache()
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
