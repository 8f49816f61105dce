"""ORACLE (test infrastructure, never imported by the product) for SURVEY.md 8(f3): CPU restatement of the reference's
point-cloud metrics, /root/reference/evaluation/evaluation_pcd.py:575-588 (compute_chamfer_distance) and :591-609
(compute_fscore).  Like the reference it uses scipy.spatial.cKDTree (the reference's own dependency, scipy is in this
image) in float64; ``nn_bruteforce`` is an independent numpy cross-check of the tree for small inputs.

Pinned: tests/golden/make_golden_chamfer.py extracts the two reference function definitions with ``ast`` (the module
itself needs trimesh / matplotlib, absent here), runs them unmodified and stores inputs + outputs in
tests/golden/chamfer.npz; tests/test_oracle_chamfer.py replays them through this file.
"""
import numpy as np
from scipy.spatial import cKDTree


def nn_query(targets, queries):
    """(distance, index) of the nearest target for every query: cKDTree(targets).query(queries, k=1)."""
    d, i = cKDTree(np.asarray(targets, dtype=np.float64)).query(np.asarray(queries, dtype=np.float64), k=1)
    return d, i


def nn_bruteforce(targets, queries):
    t, q = np.asarray(targets, np.float64), np.asarray(queries, np.float64)
    diff = q[:, None, :] - t[None, :, :]
    d2 = (diff[..., 0] * diff[..., 0] + diff[..., 1] * diff[..., 1]) + diff[..., 2] * diff[..., 2]
    i = d2.argmin(axis=1)   # first minimum = smallest index, the product's tie rule
    return np.sqrt(d2[np.arange(len(q)), i]), i


def chamfer_distance(points1, points2):
    """evaluation_pcd.py:575-588."""
    d_2to1, _ = nn_query(points1, points2)
    d_1to2, _ = nn_query(points2, points1)
    return np.mean(d_2to1) + np.mean(d_1to2)


def fscore(points1, points2, threshold=0.02):
    """evaluation_pcd.py:591-609 -> (fscore, precision, recall)."""
    d_2to1, _ = nn_query(points1, points2)
    d_1to2, _ = nn_query(points2, points1)
    precision, recall = np.mean(d_2to1 < threshold), np.mean(d_1to2 < threshold)
    f = 0.0 if precision + recall == 0 else 2 * precision * recall / (precision + recall)
    return f, precision, recall
