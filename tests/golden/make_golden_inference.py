"""Golden fixtures for SURVEY.md 8(f2), produced by the reference's OWN functions.

The reference modules cannot be imported here (bpy / matplotlib / omegaconf / imageio missing, and
scripts/inference_with_video_mesh.py runs init_config() at import), so the two function definitions are extracted from
the reference source files with `ast` and executed unmodified:
  * run_model_inference   (/root/reference/scripts/inference_with_video_mesh.py:132-256) with a fake model whose output
    encodes (window call index, frame position), so that the stitched result reveals the exact windowing;
  * smooth_trajectories   (/root/reference/utils/inference_utils.py:99-195) with method='combined'.

    python tests/golden/make_golden_inference.py
"""
import ast
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def extract(path, name, extra_globals, also=()):
    src = open(path).read()
    tree = ast.parse(src)
    nodes = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in (name,) + tuple(also)]
    ns = dict(extra_globals)
    exec(compile(ast.Module(nodes, []), path, "exec"), ns)
    return ns[name]


class Cfg(dict):
    __getattr__ = dict.__getitem__


def main():
    out_dir = os.path.dirname(os.path.abspath(__file__))
    from scipy.ndimage import gaussian_filter1d
    from scipy.signal import savgol_filter
    smooth = extract(os.path.join(REF, "utils/inference_utils.py"), "smooth_trajectories",
                     dict(torch=torch, np=np, gaussian_filter1d=gaussian_filter1d, savgol_filter=savgol_filter, print=lambda *a, **k: None),
                     also=("OneEuroFilter",))
    run_inf = extract(os.path.join(REF, "scripts/inference_with_video_mesh.py"), "run_model_inference",
                      dict(torch=torch, np=np, print=lambda *a, **k: None))

    g = torch.Generator().manual_seed(7)
    # smoothing: random walk with many sub-threshold steps
    B, T, N = 1, 23, 157
    steps = torch.randn(B, T, N, 3, generator=g) * 0.004
    steps[torch.rand(B, T, N, generator=g) < 0.5] *= 0.1
    trajs = torch.cumsum(steps, dim=1) + torch.rand(B, 1, N, 3, generator=g)
    sm = smooth(trajs, method="combined", motion_threshold=0.002, window_size=3, sigma=1.0)
    # the other methods of the same function (NumPy 2.x scalar rules: float32 scalars stay float32 against Python floats)
    small = trajs[:, :, :40].contiguous()
    extra = {}
    for w, po in ((3, 2), (5, 2), (7, 3), (4, 2)):
        extra[f"savgol_w{w}_p{po}"] = smooth(small, method="savgol", window_size=w, savgol_polyorder=po).numpy()
    extra["oneeuro_default"] = smooth(small, method="oneeuro", oneeuro_mincutoff=1.0, oneeuro_beta=0.007).numpy()
    extra["oneeuro_b05"] = smooth(small, method="oneeuro", oneeuro_mincutoff=0.3, oneeuro_beta=0.5).numpy()
    extra["gaussian_s15"] = smooth(small, method="gaussian", sigma=1.5).numpy()
    np.savez_compressed(os.path.join(out_dir, "inference_smooth.npz"), trajs=trajs.numpy(), smoothed=sm.numpy(), small=small.numpy(),
                        numpy_version=np.array(np.__version__), **extra)

    # windowing: fake model tags each output frame with 1000 * call_index + position-in-window, and records its input frames
    cases = {}
    for total_T, chunk in [(5, 8), (8, 8), (9, 8), (15, 8), (16, 8), (22, 8), (23, 8), (29, 8), (12, 12), (13, 12), (40, 12), (7, 4)]:
        calls = []

        def model(sample):
            idx = len(calls)
            frames = sample["rgb_video"][0, :, 0, 0, 0].tolist()     # frame ids were written into the pixels
            calls.append(frames)
            Tn = len(frames)
            out = torch.zeros(1, Tn, 3, 3)
            for p in range(Tn):
                out[0, p] = 1000 * idx + p
            return {"pcd_moved": out}

        video = torch.arange(total_T, dtype=torch.float32).view(total_T, 1, 1, 1).expand(total_T, 2, 2, 3).contiguous()
        cfg = Cfg(training=Cfg(frames=chunk, use_amp=False, amp_dtype="bf16"))
        cfg.training.get = lambda k, d=None, _t=cfg.training: dict.get(_t, k, d)
        inp = {"ref_pcd": torch.full((1, 3, 3), -1.0)}
        tr = run_inf(model, inp, video, cfg, "cpu")
        cases[f"T{total_T}_c{chunk}_calls"] = np.array([np.array(c, dtype=np.int64) for c in calls], dtype=object)
        cases[f"T{total_T}_c{chunk}_out"] = tr[0, :, 0, 0].numpy()
    np.savez_compressed(os.path.join(out_dir, "inference_windows.npz"), **cases, allow_pickle=True)
    print("ok", {k: v.shape for k, v in cases.items() if k.endswith("_out")})


if __name__ == "__main__":
    main()
