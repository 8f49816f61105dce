"""The inference steps either side of the hot path, same call surface as the reference:

 * ``run_model_inference(model, input_data, video_tensor, config, device)`` -- the sliding-window scheduler of
   /root/reference/scripts/inference_with_video_mesh.py:132-256 (windows of ``training.frames`` frames, stride chunk-1, frame 0
   prepended as anchor, right-aligned last window, stitched with frame 0 := ref_pcd).  Windows are independent model calls:
   with ``world_size > 1`` they are dealt round-robin to ranks and all-gathered (clips shard).
 * ``smooth_trajectories(trajs, method=...)`` -- /root/reference/utils/inference_utils.py:99-145 on the GPU
   (``m324_smooth_trajectories``), methods 'threshold', 'gaussian', 'combined' (what the shipped scripts use).
"""
import torch

from . import ops


def window_plan(total_T, chunk):
    """(start index, frame indices) per window -- inference_with_video_mesh.py:176-194."""
    if total_T <= chunk:
        return [(0, list(range(total_T)))]
    slide = chunk - 1
    starts = list(range(0, total_T - chunk + 1, slide))
    if starts and (starts[-1] + chunk < total_T):
        starts.append(total_T - chunk)
    return [(s, list(range(chunk)) if i == 0 else [0] + list(range(s + 1, s + chunk))) for i, s in enumerate(starts)]


def _stitch(outs, starts, ref_pcd):
    """inference_with_video_mesh.py:219-251."""
    n = len(outs)
    if n == 0:
        return None
    if len(starts) < 2:
        t = outs[0].clone()
        t[:, 0] = ref_pcd
        return t
    merged = []
    for i in range(n):
        if i == 0 and i != n - 2:
            c = outs[i].clone()
            c[:, 0] = ref_pcd
            merged.append(c)
        elif i < n - 2:
            merged.append(outs[i][:, 1:])
        elif i == n - 2:
            keep = max(starts[-1] - starts[-2], 0)
            if keep > 0 and n != 2:
                merged.append(outs[i][:, 1:1 + keep])
            elif keep > 0 and i == 0 and n == 2:
                c = outs[i].clone()
                c[:, 0] = ref_pcd
                merged.append(c[:, :1 + keep])
        elif i == n - 1:
            merged.append(outs[i][:, 1:])
    return torch.cat(merged, dim=1) if merged else None


def run_model_inference(model, input_data, video_tensor, config, device, rank=0, world_size=1):
    """Returns trajectories [1, total_T, N, 3] (or None), like the reference function."""
    tr = config.training
    chunk = tr.get("frames", 12) if hasattr(tr, "get") else getattr(tr, "frames", 12)
    total_T = video_tensor.shape[0]
    plan = window_plan(total_T, chunk)
    if total_T <= chunk:
        sample = dict(input_data)
        sample["rgb_video"] = video_tensor[None].float().to(device)
        out = model(sample)
        return out["pcd_moved"].float() if isinstance(out, dict) and "pcd_moved" in out else None
    outs = [None] * len(plan)
    for i, (start, frames) in enumerate(plan):
        if i % world_size != rank:
            continue
        sample = dict(input_data)
        sample["rgb_video"] = video_tensor[frames][None].float().to(device)
        out = model(sample)
        if isinstance(out, dict) and "pcd_moved" in out:
            outs[i] = out["pcd_moved"].float().clone()   # the model reuses its output workspace between calls
    if world_size > 1:
        import torch.distributed as dist
        shape = next(o for o in outs if o is not None).shape
        for i in range(len(plan)):
            buf = outs[i] if outs[i] is not None else torch.empty(shape, device=device)
            dist.broadcast(buf, src=i % world_size)
            outs[i] = buf
    outs = [o for o in outs if o is not None]
    return _stitch(outs, [s for s, _ in plan], input_data["ref_pcd"])


def smooth_trajectories(trajs, method="combined", motion_threshold=0.005, window_size=3, sigma=1.0, savgol_polyorder=2,
                        oneeuro_mincutoff=1.0, oneeuro_beta=0.007, visualization_dir=None):
    """Same signature as the reference.  trajs: [B, T, N, 3] CUDA fp32."""
    if method not in ("threshold", "gaussian", "combined"):
        raise NotImplementedError(f"smooth_trajectories(method={method!r}) is not built on the GPU path "
                                  "(the shipped inference scripts use 'combined')")
    if not trajs.is_cuda:
        raise RuntimeError("smooth_trajectories (libm324) runs on CUDA tensors only")
    x = trajs.detach().float().contiguous()
    out = torch.empty_like(x)
    ops.smooth_trajectories(x, out, motion_threshold, sigma, method in ("threshold", "combined"), method in ("gaussian", "combined"))
    return out.to(trajs.dtype)
