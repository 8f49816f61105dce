"""Chamfer / F-score kernel timing at the reference's evaluation size (50 000 x 50 000 points per frame) next to the
reference's cKDTree path on the host cores."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from motion324_b200.evaluation import evaluation_pcd as ev  # noqa: E402
from oracle import chamfer_oracle as co  # noqa: E402

n, F = 50000, int(os.environ.get("M324_F", "8"))
rng = np.random.default_rng(0)
p1 = rng.uniform(-0.5, 0.5, size=(F, n, 3))
p2 = p1[:, rng.permutation(n)] + rng.normal(size=(F, n, 3)) * 0.01
for dtype in (torch.float64, torch.float32):
    a, b = torch.from_numpy(p1).to("cuda", dtype), torch.from_numpy(p2).to("cuda", dtype)
    for _ in range(2):
        out = ev.chamfer_fscore_batch(a, b)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(3):
        out = ev.chamfer_fscore_batch(a, b)
    e.record()
    e.synchronize()
    ms = s.elapsed_time(e) / 3
    pairs = 2.0 * F * n * n
    print(f"GPU {dtype}: {ms:.2f} ms for {F} frames = {ms / F:.3f} ms/frame, {pairs / ms / 1e6:.1f} G pair/s, {pairs * 9 / ms / 1e9:.2f} TFLOP/s fp64-pipe ops")
t0 = time.perf_counter()
cd = co.chamfer_distance(p1[0], p2[0]); fs = co.fscore(p1[0], p2[0])
dt = time.perf_counter() - t0
print(f"CPU cKDTree (reference path, 1 thread): {dt * 1e3:.1f} ms/frame; chamfer {cd:.6f} vs GPU {float(out[0, 0]):.6f}; fscore {fs[0]:.4f} vs {float(out[0, 1]):.4f}")
