"""GPU drop-ins for the point-cloud metrics of the reference's evaluation script:
``compute_chamfer_distance`` (/root/reference/evaluation/evaluation_pcd.py:575-588) and ``compute_fscore`` (:591-609),
same names, argument meaning and return values (Python floats), plus a batched entry point that evaluates all frames of
a sequence in two kernel launches instead of two cKDTree builds + queries per frame (:844-892).

All arithmetic runs in libm324.so (m324_chamfer_nn / m324_chamfer_reduce, include/m324.h): exact brute-force nearest
neighbours in float64, the reference's dtype.  No CPU / scipy fallback: without a CUDA device or the library this raises.
"""
import numpy as np
import torch

from .. import ops


def _as_cuda_points(p, device):
    """numpy / torch, [n, 3] or [F, n, 3], fp32 or fp64 -> contiguous CUDA tensor [F, n, 3] (dtype kept; others -> fp64)."""
    t = torch.from_numpy(np.ascontiguousarray(p)) if isinstance(p, np.ndarray) else p
    if not torch.is_tensor(t):
        t = torch.as_tensor(np.asarray(p))
    if t.dtype not in (torch.float32, torch.float64):
        t = t.double()
    if t.dim() == 2:
        t = t.unsqueeze(0)
    if t.dim() != 3 or t.shape[-1] != 3 or t.shape[1] == 0:
        raise ValueError(f"expected points of shape [n, 3] or [F, n, 3] with n > 0, got {tuple(t.shape)}")
    return t.to(device).contiguous()


def nearest_neighbours(points1, points2, return_indices=True, device=None):
    """Both directions at once.  Returns dict(dist1, idx1, dist2, idx2) of CUDA tensors:
    dist1/idx1 [F, n2] = cKDTree(points1).query(points2, k=1); dist2/idx2 [F, n1] = cKDTree(points2).query(points1, k=1)."""
    if not torch.cuda.is_available():
        raise RuntimeError("motion324_b200.evaluation runs on a CUDA device only (no CPU fallback)")
    device = torch.device(device) if device is not None else (
        points1.device if torch.is_tensor(points1) and points1.is_cuda else torch.device("cuda", torch.cuda.current_device()))
    a, b = _as_cuda_points(points1, device), _as_cuda_points(points2, device)
    if a.dtype != b.dtype:
        a, b = a.double(), b.double()
    if a.shape[0] != b.shape[0]:
        raise ValueError(f"frame counts differ: {a.shape[0]} vs {b.shape[0]}")
    F, n1, n2 = a.shape[0], a.shape[1], b.shape[1]
    with torch.cuda.device(device):
        dist1 = torch.empty(F, n2, device=device, dtype=torch.float64)
        dist2 = torch.empty(F, n1, device=device, dtype=torch.float64)
        idx1 = torch.empty(F, n2, device=device, dtype=torch.int32) if return_indices else None
        idx2 = torch.empty(F, n1, device=device, dtype=torch.int32) if return_indices else None
        ops.chamfer_nn(a, b, dist1, idx1, dist2, idx2)
    return dict(dist1=dist1, idx1=idx1, dist2=dist2, idx2=idx2)


def chamfer_fscore_batch(points1, points2, threshold=0.02, device=None):
    """[F, n1, 3] x [F, n2, 3] -> float64 CUDA tensor [F, 4] = (chamfer, fscore, precision, recall) per frame."""
    nn = nearest_neighbours(points1, points2, return_indices=False, device=device)
    out = torch.empty(nn["dist1"].shape[0], 4, device=nn["dist1"].device, dtype=torch.float64)
    with torch.cuda.device(out.device):
        ops.chamfer_reduce(nn["dist1"], nn["dist2"], threshold, out)
    return out


def compute_chamfer_distance(points1, points2):
    """evaluation_pcd.py:575-588: mean NN distance points2 -> points1 plus mean NN distance points1 -> points2."""
    return float(chamfer_fscore_batch(points1, points2)[0, 0])


def compute_fscore(points1, points2, threshold=0.02):
    """evaluation_pcd.py:591-609: harmonic mean of precision (dist1 < threshold) and recall (dist2 < threshold)."""
    return float(chamfer_fscore_batch(points1, points2, threshold)[0, 1])
