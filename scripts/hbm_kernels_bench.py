"""The HBM-bound kernels of the path and of its neighbours (SURVEY.md 8(d) "HBM bounds ...", 8(f2)-(f4)) at the benchmark's shapes:
CUDA-event timing (L2 flushed before every launch), ALGORITHMIC bytes per launch (stated per kernel below),
achieved GB/s against the measured HBM copy peak (MEASURED_PEAKS.json), and -- for the f2 / f3 / f4 rows -- the reference's own
CPU code path timed beside it on the host cores.

    python scripts/hbm_kernels_bench.py                # JSON lines
    ncu --set full --clock-control none -k regex:'layernorm|head3|preprocess|assemble|smooth|track_points|sample_' ... python scripts/hbm_kernels_bench.py --once
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from motion324_b200 import ops  # noqa: E402
from motion324_b200.inference import smooth_trajectories  # noqa: E402
from motion324_b200.dataset import dataset_utils as du  # noqa: E402
from motion324_b200.evaluation import evaluation_pcd as ev  # noqa: E402

once = "--once" in sys.argv
dev = torch.device("cuda")
peak = 6451.8
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", peak)
except Exception:
    pass
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, iters=10):
    if once:
        fn(); torch.cuda.synchronize()
        return None
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); e.synchronize()
        tot += s.elapsed_time(e)
    return tot / iters


def report(name, ms, nbytes, what, **extra):
    if ms is None:
        return
    gbs = nbytes / (ms * 1e-3) / 1e9
    print(json.dumps(dict(kernel=name, ms=ms, algorithmic_bytes=nbytes, bytes_are=what, achieved_gbs=gbs, hbm_peak_gbs_measured=peak,
                          frac_of_hbm_peak=gbs / peak, **extra)), flush=True)


d = 768
g = torch.Generator(device="cpu").manual_seed(0)
# ---- layernorm_kernel: trunk rows (32 x 324) and one decoder chunk (32 x 4096 rows): read fp32 row, write fp16 row
for rows in (32 * 324, 32 * 4096):
    x = torch.randn(rows, d, generator=g).to(dev)
    w = torch.ones(d, device=dev)
    h = torch.empty(rows, d, device=dev, dtype=torch.float16)
    ms = timed(lambda: ops.layernorm(x, w, None, 1e-5, rows, d, out16=h, ldo16=d))
    report(f"layernorm_kernel rows={rows}", ms, rows * d * (4 + 2), "fp32 row read + fp16 row written")
    del x, h
# ---- head3_mse_kernel: 131072 rows x 768 fp32 read, 12 B / row written, target read
rows = 32 * 4096
hb = torch.randn(rows, d, generator=g).to(dev)
w3, b3 = (torch.randn(3, d, generator=g) * 0.02).to(dev), torch.zeros(3, device=dev)
out, tgt = torch.empty(rows, 3, device=dev), torch.randn(rows, 3, generator=g).to(dev)
part = torch.empty(1 << 16, device=dev)
ms = timed(lambda: ops.head3_mse(hb, d, w3, b3, rows, d, out, tgt, part))
report("head3_mse_kernel rows=131072", ms, rows * (d * 4 + 12 + 12), "fp32 hidden row read + 3 outputs written + 3 targets read")
del hb
# ---- preprocess_kernel: 32 frames 224x224x3 fp32 read, im2col fp16 [32*256, 640] written
F = 32
video = torch.rand(F, 224, 224, 3, generator=g).to(dev)
patches = torch.empty(F * 256, 640, device=dev, dtype=torch.float16)
ms = timed(lambda: ops.preprocess_frames(video, F, 224, 224, 224, patches, 640, 640))
report("preprocess_kernel 32 frames", ms, F * 224 * 224 * 3 * 4 + F * 256 * 640 * 2, "fp32 frames read + fp16 im2col patches written")
# ---- assemble_tokens_kernel: 32 x 324 tokens x 768 fp32 written, 256/324 of them read from the DINO stream (+ pos_embed)
T, M, npatch = 32, 64, 256
xd = torch.randn(T * 257, d, generator=g).to(dev)
pos = torch.randn(T * npatch, d, generator=g).to(dev)
sp = torch.randn(4, d, generator=g).to(dev)
mesh = torch.randn(M, d, generator=g).to(dev)
xo = torch.empty(T * 324, d, device=dev)
ones, zeros = torch.ones(d, device=dev), torch.zeros(d, device=dev)
ms = timed(lambda: ops.assemble_tokens(xd, ones, zeros, 1e-6, pos, sp, sp, mesh, ones, 1e-5, 1, T, M, npatch, d, xo))
report("assemble_tokens_kernel 32 frames", ms, T * 324 * d * 4 + 2 * T * npatch * d * 4, "trunk input written + DINO rows and pos_embed read")
# ---- f2: smooth_kernel at the chili size (159 frames x 13465 vertices) vs the reference's per-vertex scipy loop
Tt, Nv = 159, 13465
trajs = (torch.randn(1, Tt, Nv, 3, generator=g).cumsum(1) * 0.003).to(dev)
ms = timed(lambda: smooth_trajectories(trajs, method="combined", motion_threshold=0.002, sigma=1.0))
cpu_ms = None
if not once:
    from oracle import inference_oracle as io
    tc = trajs.cpu()
    t0 = time.perf_counter(); io.smooth_trajectories(tc, 0.002, 1.0); cpu_ms = (time.perf_counter() - t0) * 1e3
report("smooth_kernel 159x13465 (f2)", ms, 2 * Tt * Nv * 3 * 4, "trajectories read + written", cpu_ms_reference_path=cpu_ms,
       cpu_note="oracle restatement of utils/inference_utils.py:123-145 (vectorised scipy gaussian_filter1d: FASTER than the reference's N*3 Python-level calls)")
# ---- f4: track_points_kernel + sample_texture_kernel at config (c) clip size (12 frames, 4096 samples, 20k-vertex mesh)
rng = np.random.default_rng(0)
V, Fc, S, Tf = 20000, 40000, 4096, 12
verts = rng.normal(size=(Tf, V, 3))
faces = rng.integers(0, V, size=(Fc, 3))
fidx = rng.integers(0, Fc, size=S)
bary = rng.dirichlet(np.ones(3), size=S)
fuv = rng.random(size=(Fc, 3, 2))
tex = rng.integers(0, 255, size=(1024, 1024, 3), dtype=np.uint8)
vn = rng.normal(size=(Tf, V, 3))
args = [torch.as_tensor(a).to(dev) for a in (verts, faces, fidx, bary, fuv, tex, vn)]
ms = timed(lambda: du.track_with_normal_rgb(args[0], args[1], args[2], args[3], args[4], args[5], vertex_normals=args[6]))
cpu_ms = None
if not once:
    t0 = time.perf_counter()
    tri = verts[:, faces[fidx]]                                            # the NumPy gathers of dataset_utils.py:100-127
    pts = (tri * bary[None, :, :, None]).sum(2)
    nr = (vn[:, faces[fidx]] * bary[None, :, :, None]).sum(2)
    nr /= np.linalg.norm(nr, axis=-1, keepdims=True)
    cpu_ms = (time.perf_counter() - t0) * 1e3
report("track_points + sample_texture (f4) 12 frames x 4096 samples", ms, Tf * S * (3 * 3 * 8 * 2 + 2 * 12) + S * (24 + 8 + 48 + 3 + 12),
       "vertex + normal gathers (fp64) read, points + normals written, uv / texel gathers", cpu_ms_numpy=cpu_ms, includes="host-side tensor conversion of the wrapper")
# ---- f3: Chamfer / F-score, 50k x 50k per frame (fp64-pipe bound, not HBM): ms per frame beside scipy cKDTree
n, Ff = 50000, 4
p1 = rng.uniform(-0.5, 0.5, size=(Ff, n, 3))
p2 = p1[:, rng.permutation(n)] + rng.normal(size=(Ff, n, 3)) * 0.01
a, b = torch.from_numpy(p1).to(dev), torch.from_numpy(p2).to(dev)
ms = timed(lambda: ev.chamfer_fscore_batch(a, b), iters=3)
if ms is not None:
    from oracle import chamfer_oracle as co
    t0 = time.perf_counter(); co.chamfer_distance(p1[0], p2[0]); co.fscore(p1[0], p2[0]); cpu = (time.perf_counter() - t0) * 1e3
    print(json.dumps(dict(kernel="nn_kernel + chamfer_reduce_kernel (f3) 50k x 50k fp64", ms_per_frame=ms / Ff, pair_distances_per_s=2.0 * Ff * n * n / (ms * 1e-3),
                          bound="fp64 pipe (exact brute force, 9 flop / pair)", cpu_ms_per_frame_reference_ckdtree=cpu)), flush=True)
print("done")
