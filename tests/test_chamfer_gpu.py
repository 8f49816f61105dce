"""GPU parity of the point-cloud metrics through the C ABI (m324_chamfer_nn / m324_chamfer_reduce) against
the oracle (cKDTree restatement pinned to the reference's functions) and the committed golden values.
Bar: neighbour INDICES bit-exact; distances and metrics to float64 rounding (rtol 1e-12, stated below)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden", "chamfer.npz")
RTOL = 1e-12


@pytest.fixture(scope="module")
def ev():
    from motion324_b200.evaluation import evaluation_pcd
    return evaluation_pcd


def test_golden_reference_values(ev):
    g = np.load(GOLD)
    for k in range(5):
        p1, p2 = g[f"c{k}_p1"], g[f"c{k}_p2"]
        assert ev.compute_chamfer_distance(p1, p2) == pytest.approx(float(g[f"c{k}_chamfer"]), rel=RTOL)
        for t, ref in zip((0.02, 0.05, 0.001), g[f"c{k}_fscore"]):
            assert ev.compute_fscore(p1, p2, threshold=t) == pytest.approx(float(ref), rel=RTOL, abs=0)


@pytest.mark.parametrize("n1,n2,dtype", [(1, 1, np.float64), (5, 1500, np.float64), (1024, 1024, np.float64), (1025, 2049, np.float32),
                                         (3000, 777, np.float32)])
def test_indices_and_distances_vs_oracle(ev, n1, n2, dtype):
    from oracle import chamfer_oracle as co
    rng = np.random.default_rng(n1 * 7 + n2)
    F = 3
    p1 = rng.uniform(-0.5, 0.5, size=(F, n1, 3)).astype(dtype)
    p2 = (rng.uniform(-0.5, 0.5, size=(F, n2, 3)) * 1.1).astype(dtype)
    nn = ev.nearest_neighbours(p1, p2)
    met = ev.chamfer_fscore_batch(p1, p2, threshold=0.05).cpu().numpy()
    for f in range(F):
        d1, i1 = co.nn_query(p1[f], p2[f])
        d2, i2 = co.nn_query(p2[f], p1[f])
        assert np.array_equal(nn["idx1"][f].cpu().numpy(), i1.astype(np.int32))      # bit-exact indices
        assert np.array_equal(nn["idx2"][f].cpu().numpy(), i2.astype(np.int32))
        assert np.allclose(nn["dist1"][f].cpu().numpy(), d1, rtol=RTOL, atol=0)
        assert np.allclose(nn["dist2"][f].cpu().numpy(), d2, rtol=RTOL, atol=0)
        fs, pr, rc = co.fscore(p1[f], p2[f], 0.05)
        assert np.allclose(met[f], [co.chamfer_distance(p1[f], p2[f]), fs, pr, rc], rtol=RTOL, atol=0)


def test_ties_take_smallest_index_and_duplicates(ev):
    p1 = np.array([[0.0, 0, 0], [1.0, 0, 0], [0.0, 0, 0], [1.0, 0, 0]])     # duplicated targets
    p2 = np.array([[0.5, 0, 0], [0.0, 0, 0], [0.9, 0, 0]])
    nn = ev.nearest_neighbours(p1, p2)
    assert nn["idx1"][0].tolist() == [0, 0, 1]
    assert nn["dist1"][0].tolist() == [0.5, 0.0, pytest.approx(0.1, rel=1e-15)]
    assert ev.compute_chamfer_distance(p1, p1) == 0.0 and ev.compute_fscore(p1, p1) == 1.0


def test_full_size_properties(ev):
    """50 000 x 50 000 points per frame (the reference's num_samples), 2 frames: size-independent properties instead of the
    tree -- symmetry of the Chamfer distance under swapping the clouds (precision <-> recall), identity, and a planted
    neighbour structure whose answer is known in closed form."""
    rng = np.random.default_rng(5)
    n = 50000
    p1 = rng.uniform(-0.5, 0.5, size=(2, n, 3))
    perm = np.stack([rng.permutation(n) for _ in range(2)])
    shift = np.array([1e-4, -2e-4, 3e-4])
    p2 = np.stack([p1[f][perm[f]] for f in range(2)]) + shift            # every point has a partner at |shift|
    a = ev.chamfer_fscore_batch(p1, p2, threshold=0.02).cpu().numpy()
    b = ev.chamfer_fscore_batch(p2, p1, threshold=0.02).cpu().numpy()
    assert np.allclose(a[:, 0], b[:, 0], rtol=RTOL) and np.allclose(a[:, 2], b[:, 3], rtol=0, atol=0) and np.allclose(a[:, 1], b[:, 1], rtol=RTOL)
    nn = ev.nearest_neighbours(p1, p2)
    inv = np.empty_like(perm)
    for f in range(2):
        inv[f, perm[f]] = np.arange(n)
    d = np.linalg.norm(shift)
    # the planted partner is the nearest neighbour unless another point lies within |shift| (probability ~ n * 4/3 pi d^3 ~ 1e-5)
    assert (nn["idx1"].cpu().numpy() == perm).mean() > 0.9999 and (nn["idx2"].cpu().numpy() == inv).mean() > 0.9999
    assert np.all(nn["dist1"].cpu().numpy() <= d * (1 + 1e-9)) and np.all(nn["dist2"].cpu().numpy() <= d * (1 + 1e-9))
    assert a[:, 0] == pytest.approx(2 * d, rel=1e-3) and np.all(a[:, 1] == 1.0)
    same = ev.chamfer_fscore_batch(p1, p1).cpu().numpy()
    assert np.all(same[:, 0] == 0.0) and np.all(same[:, 1:] == 1.0)


def test_errors(ev):
    with pytest.raises(ValueError):
        ev.compute_chamfer_distance(np.zeros((0, 3)), np.zeros((4, 3)))
    with pytest.raises(ValueError):
        ev.compute_chamfer_distance(np.zeros((4, 2)), np.zeros((4, 3)))
    with pytest.raises(ValueError):
        ev.chamfer_fscore_batch(np.zeros((2, 4, 3)), np.zeros((3, 4, 3)))
