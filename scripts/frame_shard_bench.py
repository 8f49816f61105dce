"""One long clip, frames sharded over the ranks (Motion_Latent_Model.frame_parallel): strong scaling of the long-sequence
sweep (BASELINE.json configs[4]).  Launch with torchrun; world size 1 = the unsharded forward.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/frame_shard_bench.py --frames 128 256 --points 4096
"""
import argparse, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from motion324_b200.model.Pcd_motion import Motion_Latent_Model
from motion324_b200.utils.config import make_config
from motion324_b200.utils import synthetic as orc  # seeded weights / inputs generator

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, nargs="+", default=[128])
ap.add_argument("--points", type=int, default=4096)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--overlap", action="store_true", help="own keys first while the K|V all-gather travels (partial launches + merge)")
a = ap.parse_args()
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.distributed.init_process_group("nccl", device_id=dev)
for T in a.frames:
    model = Motion_Latent_Model(make_config(frames=T))
    model.load_state_dict(orc.init_state_dict(0, dict(frames=T)), strict=True)
    model = model.to(dev)
    model.eval()
    model.frame_parallel(world > 1)
    model.fp_overlap = a.overlap
    one = orc.make_inputs(seed=1, B=1, T=1, N=a.points, S=a.points)
    g = torch.Generator().manual_seed(2)
    sample = {k: v.to(dev) for k, v in one.items()}
    sample["rgb_video"] = torch.rand(1, T, 224, 224, 3, generator=g).to(dev)
    sample["point_clouds"] = sample["ref_pcd"][:, None] + 0.05 * torch.randn(1, T, a.points, 3, generator=g).to(dev)
    for _ in range(2):
        ret = model(sample)
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(a.steps):
        ret = model(sample)
    e.record()
    torch.cuda.synchronize()
    ms = torch.tensor([s.elapsed_time(e) / a.steps], device=dev)
    if world > 1:
        torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps(dict(what="one clip, frames sharded over ranks" if world > 1 else "one clip, one GPU", frames=T, points=a.points,
                              n_gpus=world, ms_per_clip=float(ms[0]), frames_per_s=T / float(ms[0]) * 1e3, loss=float(ret.loss_metrics.loss),
                              kv_allgather_bytes_per_layer=T * 324 * 1536 * 2 if world > 1 else 0, overlap=bool(world > 1 and a.overlap), scaling="strong")), flush=True)
    del model, sample, ret
    torch.cuda.empty_cache()
if world > 1:
    torch.distributed.destroy_process_group()
