"""BASELINE config (d): inference on examples/chili.glb + chili.mp4 through the product entry point
(motion324_b200/scripts/inference_with_video_mesh.py) against tests/golden/chili_T159.npz, which holds outputs of the
reference's own code on the same files (tests/golden/make_golden_chili.py): sample colours from the unmodified reference loop
(bytes, exact), nearest-sample indices from scipy's cKDTree (exact), trajectories from the unmodified reference model in fp32
(1e-3).  T = 159 frames with training.frames = 256 (trilinear pos-embed resize), 13,465 vertices = 4 decoder chunks in the
reference, S = 16,384, one global attention over 51,516 tokens.  The demo assets travel with the staged reference (oracle/_ref)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from oracle import build_ref  # noqa: E402
from oracle import motion324_oracle as orc  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden", "chili_T159.npz")
if not (build_ref.available() and os.path.exists(GOLD)):
    pytest.skip("the staged reference (demo assets) or the chili fixture is missing", allow_module_level=True)


def test_chili_demo_matches_the_reference_pipeline():
    from motion324_b200.model.Pcd_motion import Motion_Latent_Model
    from motion324_b200.scripts.inference_with_video_mesh import (load_video_from_path, prepare_mesh_data, run_model_inference,
                                                                  select_frames, smooth_trajectories)
    from motion324_b200.utils.config import make_config
    g = np.load(GOLD)
    root = build_ref.root()
    cfg = make_config(frames=256, num_shape_samples=16384)
    dev = torch.device("cuda")
    # ---- mesh -> tensors (f4: GLB reader, host sampler, m324_sample_albedo, m324_chamfer_nn)
    input_data, mesh, faces = prepare_mesh_data(cfg, os.path.join(root, "examples", "chili.glb"), dev)
    N, S = int(g["N"]), 16384
    assert tuple(input_data["ref_pcd"].shape) == (1, N, 3) and tuple(input_data["ref_shape_pcd"].shape) == (1, S, 3)
    assert faces.shape == (19753, 3)
    assert abs(float(input_data["ref_shape_pcd"].double().sum()) - float(g["shape_pcd_sum"])) < 1e-6      # same host sampler
    rgb8 = torch.round(input_data["ref_shape_rgbs"][0] * 255.0).to(torch.uint8).cpu().numpy()
    assert np.array_equal(rgb8, g["shape_rgb_u8"])                                # texels gathered by the kernel == the reference loop's
    assert torch.equal(input_data["ref_shape_rgbs"][0].cpu(), torch.from_numpy(g["shape_rgb_u8"].astype(np.float32) / np.float32(255.0)))
    nearest = torch.from_numpy(g["nearest"].astype(np.int64))
    assert torch.equal(input_data["ref_rgb"][0].cpu(), input_data["ref_shape_rgbs"][0].cpu()[nearest])    # NN colour transfer == cKDTree's
    # ---- video
    video = select_frames(load_video_from_path(os.path.join(root, "examples", "chili.mp4")), cfg.training.frames)
    T = int(g["T"])
    assert tuple(video.shape) == (T, 720, 720, 3)
    assert abs(float(video.double().sum()) - float(g["video_sum"])) < 1e-3
    # ---- model (one call: total_T <= chunk), weights = the seeded random init of a 256-frame model
    model = Motion_Latent_Model(cfg)
    model.load_state_dict(orc.init_state_dict(seed=0, cfg=dict(frames=256)), strict=True)
    model = model.to(dev)
    model.eval()
    input_data["rgb_video"] = video
    trajs = run_model_inference(model, input_data, video, cfg, dev)
    torch.cuda.synchronize()
    assert tuple(trajs.shape) == (1, T, N, 3)
    stride = int(g["vertex_stride"])
    got, ref = trajs.cpu(), torch.from_numpy(g["pcd_moved"])
    err = orc.rel_l2(got[:, :, ::stride], ref)
    worst = max(orc.rel_l2(got[:, t, ::stride], ref[:, t]) for t in range(T))
    print(f"chili: pcd_moved rel-L2 {err:.3e} (worst frame {worst:.3e})")
    assert err < 1e-3 and worst < 2e-3, (err, worst)
    assert abs(float(got.double().pow(2).sum()) - float(g["pcd_moved_sqsum"])) < 2e-3 * float(g["pcd_moved_sqsum"])
    # ---- post-processing runs on the result (parity of the smoothing itself: test_smooth_trajectories_matches_reference_golden)
    sm = smooth_trajectories(trajs, method="combined", motion_threshold=0.002, sigma=1.0)
    assert tuple(sm.shape) == tuple(trajs.shape) and torch.isfinite(sm).all()
