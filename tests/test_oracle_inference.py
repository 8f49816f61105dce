"""Pin oracle/inference_oracle.py against the outputs of the reference's own run_model_inference /
smooth_trajectories (tests/golden/inference_*.npz, made by tests/golden/make_golden_inference.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import inference_oracle as io

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _fake_outs(plan):
    outs = []
    for idx, (_, frames) in enumerate(plan):
        o = torch.zeros(1, len(frames), 3, 3)
        for p in range(len(frames)):
            o[0, p] = 1000 * idx + p
        outs.append(o)
    return outs


def test_window_plan_and_stitch_match_reference():
    g = np.load(os.path.join(GOLD, "inference_windows.npz"), allow_pickle=True)
    names = sorted({k.rsplit("_", 1)[0] for k in g.files})
    assert len(names) == 12
    for name in names:
        total_T, chunk = int(name.split("_")[0][1:]), int(name.split("_")[1][1:])
        plan = io.window_plan(total_T, chunk)
        calls = [list(map(int, c)) for c in g[name + "_calls"]]
        assert [f for _, f in plan] == calls, name                      # same windows, same anchor frame
        st = io.stitch(_fake_outs(plan), [s for s, _ in plan], torch.full((1, 3, 3), -1.0), single_pass=total_T <= chunk)
        assert np.array_equal(st[0, :, 0, 0].numpy(), g[name + "_out"]), name   # same stitching, exact
        assert st.shape[1] == total_T


def test_smoothing_matches_reference():
    g = np.load(os.path.join(GOLD, "inference_smooth.npz"))
    out = io.smooth_trajectories(torch.from_numpy(g["trajs"]), motion_threshold=0.002, sigma=1.0)
    assert np.array_equal(out.numpy(), g["smoothed"])                  # same scipy routine: bit-exact
    held = (torch.from_numpy(g["trajs"])[:, 1:] - torch.from_numpy(g["trajs"])[:, :-1]).norm(dim=-1) < 0.002
    assert 0.2 < float(held.float().mean()) < 0.8                      # the fixture exercises both branches


def test_oracle_savgol_and_oneeuro_match_the_reference_function():
    """The remaining methods of smooth_trajectories (utils/inference_utils.py:148-175): the oracle against outputs of the
    reference's own function (tests/golden/make_golden_inference.py, NumPy 2 scalar rules)."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "inference_smooth.npz"))
    small = torch.from_numpy(g["small"])
    for w, po in ((3, 2), (5, 2), (7, 3), (4, 2)):
        got = io.smooth_trajectories(small, method="savgol", window_size=w, savgol_polyorder=po)
        assert np.array_equal(got.numpy(), g[f"savgol_w{w}_p{po}"]), (w, po)
    assert np.array_equal(io.smooth_trajectories(small, method="oneeuro").numpy(), g["oneeuro_default"])
    assert np.array_equal(io.smooth_trajectories(small, method="oneeuro", oneeuro_mincutoff=0.3, oneeuro_beta=0.5).numpy(), g["oneeuro_b05"])
    assert np.array_equal(io.smooth_trajectories(small, sigma=1.5, method="gaussian").numpy(), g["gaussian_s15"])
