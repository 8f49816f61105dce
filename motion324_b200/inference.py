"""The inference steps either side of the hot path, same call surface as the reference:

 * ``run_model_inference(model, input_data, video_tensor, config, device)`` -- the sliding-window scheduler of
   /root/reference/scripts/inference_with_video_mesh.py:132-256 (windows of ``training.frames`` frames, stride chunk-1, frame 0
   prepended as anchor, right-aligned last window, stitched with frame 0 := ref_pcd).  Windows are independent model calls:
   with ``world_size > 1`` they are dealt round-robin to ranks and combined by one all-reduce (clips shard).
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


def frame_sources(total_T, chunk):
    """Which window produces which output frame: a list of (window, local index, first frame, count) runs that tile
    frames 1 .. total_T-1 exactly once (frame 0 is always the reference shape).  This is the merge rule of
    inference_with_video_mesh.py:219-251 as an index plan: a window i > 0 holds the anchor at local index 0 and frames
    s_i+1 .. s_i+chunk-1 behind it; consecutive windows abut, except the right-aligned last one, which overlaps its
    predecessor and wins the overlap -- i.e. frame g comes from the LAST window whose start lies before g."""
    starts = [s for s, _ in window_plan(total_T, chunk)]
    runs = []
    for w, s in enumerate(starts):
        first = s + 1
        last = min(s + chunk - 1, total_T - 1) if w == len(starts) - 1 else min(s + chunk - 1, starts[w + 1])
        if last >= first:
            runs.append((w, first - s, first, last - first + 1))
    return runs


def run_model_inference(model, input_data, video_tensor, config, device, rank=0, world_size=1, group=None):
    """Returns trajectories [1, total_T, N, 3] (or None), like the reference function.  With world_size > 1 the windows are
    dealt round-robin to the ranks; every rank writes the frames its windows own into a zero-initialised result and ONE
    all-reduce (each frame has exactly one non-zero contributor, so the sum is exact) hands every rank the whole clip.  Ranks
    without a window (more ranks than windows) contribute zeros and still join the collective."""
    tr = config.training
    chunk = tr.get("frames", 12) if hasattr(tr, "get") else getattr(tr, "frames", 12)
    total_T = video_tensor.shape[0]
    if total_T <= chunk:
        sample = dict(input_data)
        sample["rgb_video"] = video_tensor[None].float().to(device)
        out = model(sample)
        return out["pcd_moved"].float() if isinstance(out, dict) and "pcd_moved" in out else None
    plan = window_plan(total_T, chunk)
    ref_pcd = input_data["ref_pcd"]
    trajs = torch.zeros((1, total_T) + tuple(ref_pcd.shape[1:]), device=ref_pcd.device, dtype=torch.float32)
    runs = frame_sources(total_T, chunk)
    for w, (start, frames) in enumerate(plan):
        if w % world_size != rank:
            continue
        sample = dict(input_data)
        sample["rgb_video"] = video_tensor[frames][None].float().to(device)
        out = model(sample)
        if not (isinstance(out, dict) and "pcd_moved" in out):
            raise RuntimeError("run_model_inference: the model returned no 'pcd_moved'")
        for rw, local, first, count in runs:
            if rw == w:
                trajs[:, first:first + count] = out["pcd_moved"][:, local:local + count].float()
    if world_size > 1:
        import torch.distributed as dist
        dist.all_reduce(trajs, group=group)
    trajs[:, 0] = ref_pcd.float()      # frame 0 := the reference shape (inference_with_video_mesh.py:222-224, 248-249)
    return trajs


def smooth_trajectories(trajs, method="combined", motion_threshold=0.005, window_size=3, sigma=1.0, savgol_polyorder=2,
                        oneeuro_mincutoff=1.0, oneeuro_beta=0.007, visualization_dir=None):
    """Same signature and methods as the reference ('threshold', 'gaussian', 'combined', 'savgol', 'oneeuro'; any other name
    returns a copy, like the reference's if-chain).  trajs: [B, T, N, 3] CUDA fp32."""
    if not trajs.is_cuda:
        raise RuntimeError("smooth_trajectories (libm324) runs on CUDA tensors only")
    x = trajs.detach().float().contiguous()
    out = torch.empty_like(x)
    T = x.shape[1]
    if method in ("threshold", "gaussian", "combined"):
        ops.smooth_trajectories(x, out, motion_threshold, sigma, method in ("threshold", "combined"), method in ("gaussian", "combined"))
    elif method == "savgol":
        if window_size % 2 == 0:                                   # inference_utils.py:151-152
            window_size += 1
        if T >= window_size:
            ops.filter_trajectories(x, out, 1, taps=savgol_taps(window_size, min(savgol_polyorder, window_size - 1)))
        else:
            out.copy_(x)
    elif method == "oneeuro":
        ops.filter_trajectories(x, out, 2, mincutoff=oneeuro_mincutoff, beta=oneeuro_beta)
    else:
        out.copy_(x)
    return out.to(trajs.dtype)


def savgol_taps(window_length, polyorder):
    """scipy.signal.savgol_coeffs(window_length, polyorder) (deriv 0, centred): the least-squares polynomial fit evaluated at the
    window centre, as a symmetric FIR.  Host-side constant (a (polyorder+1) x window least-squares solve in NumPy)."""
    import numpy as np
    if polyorder >= window_length:
        raise ValueError("polyorder must be less than window_length.")
    half = window_length // 2
    pos = np.arange(-half, window_length - half, dtype=np.float64)[::-1]      # scipy evaluates at x = pos - ... with reversed order
    A = pos ** np.arange(polyorder + 1).reshape(-1, 1)
    y = np.zeros(polyorder + 1)
    y[0] = 1.0
    coeffs, *_ = np.linalg.lstsq(A, y, rcond=None)
    return coeffs.tolist()
