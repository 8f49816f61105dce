"""ORACLE (test infrastructure) for SURVEY.md 8(f2): the inference steps either side of the hot path.

 * ``window_plan`` / ``stitch`` restate the sliding-window scheduler of ``run_model_inference``
   (/root/reference/scripts/inference_with_video_mesh.py:132-256): windows of ``chunk`` frames, stride chunk-1, frame 0
   prepended as anchor to every window but the first, last window right-aligned, trajectories stitched with frame 0
   overwritten by ``ref_pcd``.
 * ``smooth_trajectories`` restates method='combined' of /root/reference/utils/inference_utils.py:99-145: threshold
   filter (a point holds its previous SMOOTHED position when its RAW displacement is below ``motion_threshold``), then
   scipy.ndimage.gaussian_filter1d(sigma, mode='nearest') per vertex and axis along time.

Pinned against the reference functions themselves: tests/golden/make_golden_inference.py extracts the two function
definitions from the reference sources (the modules cannot be imported: bpy / matplotlib / omegaconf are absent), executes
them, and stores their outputs in tests/golden/inference_*.npz.
"""
import numpy as np
import torch
from scipy.ndimage import gaussian_filter1d


def window_plan(total_T, chunk):
    """inference_with_video_mesh.py:176-194 -> list of (start_idx, frame index list of the window)."""
    if total_T <= chunk:
        return [(0, list(range(total_T)))]
    slide = chunk - 1
    starts = list(range(0, total_T - chunk + 1, slide))
    if starts and (starts[-1] + chunk < total_T):
        starts.append(total_T - chunk)
    plan = []
    for i, s in enumerate(starts):
        frames = list(range(0, chunk)) if i == 0 else [0] + list(range(s + 1, s + chunk))
        plan.append((s, frames))
    return plan


def stitch(outs, starts, ref_pcd, single_pass=False):
    """inference_with_video_mesh.py:219-251.  outs: list of [1, chunk, N, 3] tensors, one per window.
    single_pass: total_T <= chunk (:157-174) -- the model output is returned as is (frame 0 NOT overwritten)."""
    if len(outs) == 0:
        return None
    if single_pass:
        return outs[0]
    if len(starts) < 2:
        t = outs[0].clone()
        t[:, 0] = ref_pcd
        return t
    merged = []
    n = len(outs)
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


def smooth_trajectories(trajs, motion_threshold=0.005, sigma=1.0, method="combined", window_size=3, savgol_polyorder=2,
                        oneeuro_mincutoff=1.0, oneeuro_beta=0.007):
    """inference_utils.py:123-175 (threshold / gaussian / combined / savgol / oneeuro); trajs [B, T, N, 3] float32 torch tensor."""
    out = trajs.clone()
    if method == "savgol":          # :148-163, vectorised over (b, n, dim): savgol_filter works along an axis
        from scipy.signal import savgol_filter
        if window_size % 2 == 0:
            window_size += 1
        if trajs.shape[1] >= window_size:
            a = savgol_filter(out.numpy(), window_length=window_size, polyorder=min(savgol_polyorder, window_size - 1), mode="nearest", axis=1)
            out = torch.from_numpy(np.ascontiguousarray(a)).to(trajs.dtype)
        return out
    if method == "oneeuro":         # :58-96, 165-175 as float32 array arithmetic (elementwise = the reference's float32 scalars, NEP 50)
        x = out.numpy()
        f32 = np.float32
        r = 2 * np.pi * 1.0 * 1.0
        alpha_d = r / (r + 1)
        res = x.copy()
        x_prev, dx_prev = x[:, 0].copy(), np.zeros_like(x[:, 0])
        for t in range(1, x.shape[1]):
            dx = x[:, t] - x_prev
            dx_hat = f32(alpha_d) * dx + f32(1 - alpha_d) * dx_prev
            cutoff = f32(oneeuro_mincutoff) + f32(oneeuro_beta) * np.abs(dx_hat)
            rr = f32(2 * np.pi) * cutoff * f32(1.0)
            alpha = rr / (rr + f32(1))
            x_hat = alpha * x[:, t] + (f32(1) - alpha) * x_prev
            res[:, t] = x_hat
            x_prev, dx_prev = x_hat, dx_hat
        return torch.from_numpy(res)
    B, T, N, _ = trajs.shape
    if method in ("threshold", "combined"):
        for b in range(B):
            for t in range(1, T):
                mask = torch.norm(trajs[b, t] - trajs[b, t - 1], dim=-1) < motion_threshold
                out[b, t][mask] = out[b, t - 1][mask]
    if method in ("gaussian", "combined"):
        a = out.cpu().numpy()
        a = gaussian_filter1d(a, sigma=sigma, axis=1, mode="nearest")   # per (b, n, dim) series along time
        out = torch.from_numpy(np.ascontiguousarray(a)).to(trajs.dtype)
    return out
