"""CPU: the data-prep oracle (oracle/dataprep_oracle.py) against outputs of the reference's own dataset_utils functions
(tests/golden/dataprep.npz, made by tests/golden/make_golden_dataprep.py).  SURVEY.md 8(f4)."""
import os

import numpy as np
import pytest

from oracle import dataprep_oracle as orc
CASES = orc.CASES

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "dataprep.npz"))


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_reference_track_with_normal_rgb(name):
    case = orc.make_case(*CASES[name])
    pts, nrm, rgb, yx = orc.track(case["vertex_frames"], case["faces"], case["face_indices"], case["bary"], case["face_uvs"],
                                  case["texture"], case["vertex_normals"])
    assert np.array_equal(pts, GOLD[name + "_points"])
    assert np.array_equal(nrm, GOLD[name + "_normals"])
    assert np.array_equal(rgb, GOLD[name + "_rgbs"])
    if name == "degenerate":
        assert np.all(nrm[:, 0] == 0.0)                       # zero-norm branch (dataset_utils.py:125) keeps the zero vector
        H, W = case["texture"].shape[:2]
        assert tuple(yx[2]) == (H - 1, W - 1)                 # u = 1, v = 0 lands exactly on the last texel


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_reference_texture_lookup(name):
    case = orc.make_case(*CASES[name])
    uvs = np.random.default_rng(CASES[name][0] + 100).uniform(-0.2, 1.2, size=(500, 2))
    texels, yx = orc.sample_texture_color_vectorized(uvs, case["texture"])
    assert np.array_equal(texels, GOLD[name + "_texels"])
    assert yx.min() >= 0 and yx[:, 0].max() <= case["texture"].shape[0] - 1
