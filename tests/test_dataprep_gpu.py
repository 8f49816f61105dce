"""GPU: m324_track_points / m324_sample_texture_colors (through the C ABI, motion324_b200/dataset/dataset_utils.py) against the
NumPy oracle and the reference's own outputs.  Integer work (face / vertex / texel gathers) is bit-exact; the float outputs are
the float64 -> float32 rounding of the NumPy result (bit-exact up to a float64 last-bit difference of the three-term sums,
checked as <= 1 float32 ulp everywhere and exact on >= 99.9 %).  SURVEY.md 8(f4)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from motion324_b200.dataset import dataset_utils as du  # noqa: E402
from oracle import dataprep_oracle as orc  # noqa: E402
CASES = orc.CASES

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "dataprep.npz"))


def _close_f32(got, ref):
    got, ref = np.asarray(got, dtype=np.float32), np.asarray(ref, dtype=np.float32)
    assert got.shape == ref.shape
    exact = np.mean(got == ref) if got.size else 1.0
    ulp = np.abs(got.view(np.int32).astype(np.int64) - ref.view(np.int32).astype(np.int64)).max() if got.size else 0
    assert exact >= 0.999 and ulp <= 1, (exact, ulp)


def _run(case, normals=True):
    return du.track_with_normal_rgb(case["vertex_frames"], case["faces"], case["face_indices"], case["bary"], case["face_uvs"],
                                    case["texture"], case["vertex_normals"] if normals else None)


@pytest.mark.parametrize("name", list(CASES))
def test_track_matches_reference_golden(name):
    case = orc.make_case(*CASES[name])
    pts, nrm, rgb, fidx = _run(case)
    assert pts.dtype == torch.float32 and pts.is_cuda and tuple(pts.shape) == GOLD[name + "_points"].shape
    _close_f32(pts.cpu().numpy(), GOLD[name + "_points"])
    _close_f32(nrm.cpu().numpy(), GOLD[name + "_normals"])
    assert np.array_equal(rgb.cpu().numpy(), GOLD[name + "_rgbs"])          # gathered bytes / 255: exact
    assert fidx is case["face_indices"]


@pytest.mark.parametrize("T,V,F,S,H,W,dtype", [(32, 13465, 26000, 16384, 1024, 1024, np.float32), (1, 3, 1, 1, 1, 1, np.float64),
                                                (4, 100, 150, 0, 4, 4, np.float32), (12, 5000, 9000, 4096, 2048, 512, np.float64)])
def test_track_matches_oracle_sizes(T, V, F, S, H, W, dtype):
    """Config-(d)-like sizes (chili.glb: 13,465 vertices, 16,384 samples), the smallest mesh, an empty sample set."""
    case = orc.make_case(7, T, V, F, S, H, W, False, dtype)
    pts, nrm, rgb, _ = _run(case)
    rp, rn, rr, yx = orc.track(case["vertex_frames"], case["faces"], case["face_indices"], case["bary"], case["face_uvs"],
                               case["texture"], case["vertex_normals"])
    if S == 0:
        assert tuple(pts.shape) == (T, 0, 3) and tuple(rgb.shape) == (T, 0, 3)
        return
    _close_f32(pts.cpu().numpy(), rp)
    _close_f32(nrm.cpu().numpy(), rn)
    assert np.array_equal(rgb.cpu().numpy(), rr)
    # the gathered texel coordinates themselves (integer work): bit-exact
    fu = torch.as_tensor(case["face_uvs"]).cuda()
    _, texel = du._sample(fu, torch.as_tensor(case["face_indices"]).cuda(), torch.as_tensor(case["bary"]).cuda(),
                          torch.as_tensor(case["texture"]).cuda())
    assert np.array_equal(texel.cpu().numpy(), yx)
    p2, n2, _, _ = _run(case, normals=False)
    assert n2 is None and torch.equal(p2, pts)


@pytest.mark.parametrize("name", list(CASES))
def test_texture_lookup_matches_reference(name):
    case = orc.make_case(*CASES[name])
    uvs = np.random.default_rng(CASES[name][0] + 100).uniform(-0.2, 1.2, size=(500, 2))
    got = du.sample_texture_color_vectorized(uvs, case["texture"])
    assert got.dtype == torch.uint8 and np.array_equal(got.cpu().numpy(), GOLD[name + "_texels"])


def test_out_of_range_indices_raise_like_numpy():
    case = orc.make_case(9, 2, 30, 40, 16, 4, 4)
    bad = dict(case, face_indices=case["face_indices"].copy())
    bad["face_indices"][3] = 40
    with pytest.raises(IndexError):
        _run(bad)
    bad = dict(case, faces=case["faces"].copy())
    bad["faces"][case["face_indices"][0], 1] = 30
    with pytest.raises(IndexError):
        _run(bad)
    with pytest.raises(RuntimeError):
        du.track_with_normal_rgb(case["vertex_frames"], case["faces"], case["face_indices"], case["bary"], case["face_uvs"],
                                 case["texture"], None, device="cpu")
