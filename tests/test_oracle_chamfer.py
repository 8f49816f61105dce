"""Pin oracle/chamfer_oracle.py against the outputs of the reference's own compute_chamfer_distance /
compute_fscore (tests/golden/chamfer.npz, made by tests/golden/make_golden_chamfer.py).  CPU only."""
import os

import numpy as np

from oracle import chamfer_oracle as co

GOLD = os.path.join(os.path.dirname(__file__), "golden", "chamfer.npz")


def test_oracle_matches_reference_functions():
    g = np.load(GOLD)
    for k in range(5):
        p1, p2 = g[f"c{k}_p1"], g[f"c{k}_p2"]
        assert co.chamfer_distance(p1, p2) == float(g[f"c{k}_chamfer"])           # same routine, same dtype: bit-exact
        for t, ref in zip((0.02, 0.05, 0.001), g[f"c{k}_fscore"]):
            assert co.fscore(p1, p2, t)[0] == float(ref)


def test_tree_agrees_with_bruteforce():
    """The kd-tree and an O(n^2) scan pick the same neighbours (no exact ties in random data) and distances to 1 ulp."""
    g = np.load(GOLD)
    for k in (1, 2, 3):
        p1, p2 = g[f"c{k}_p1"], g[f"c{k}_p2"]
        d_t, i_t = co.nn_query(p1, p2)
        d_b, i_b = co.nn_bruteforce(p1, p2)
        assert np.array_equal(i_t, i_b)
        assert np.allclose(d_t, d_b, rtol=1e-15, atol=0)
