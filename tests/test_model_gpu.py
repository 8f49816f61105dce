"""End-to-end parity of the libm324-backed Motion_Latent_Model against the oracle (fp32 CPU restatement pinned to the
reference) and against the committed golden outputs of the unmodified reference.  Run on the B200 box: pytest -m gpu.

Tolerance (BASELINE.json north_star): 1e-3 relative, measured as ||out - ref||_2 / ||ref||_2 over pcd_moved."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from motion324_b200.model.Pcd_motion import Motion_Latent_Model  # noqa: E402
from motion324_b200.utils.config import make_config  # noqa: E402
from oracle import motion324_oracle as orc  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden")
REL_TOL = 1e-3


def _build(frames):
    model = Motion_Latent_Model(make_config(frames=frames))
    model.load_state_dict(orc.init_state_dict(seed=0, cfg=dict(frames=frames)), strict=True)
    model = model.to("cuda")
    model.eval()
    return model


def _run(model, sample):
    dev = {k: v.to("cuda") for k, v in sample.items()}
    ret = model(dev)
    torch.cuda.synchronize()
    return ret


@pytest.mark.parametrize("name,c", [
    ("a_T1_N512", dict(frames=1, T=1, N=512, S=512, H=224, W=224)),
    ("resize_T3_N300", dict(frames=4, T=3, N=300, S=700, H=224, W=224)),
    ("chunk_T2_N4200", dict(frames=2, T=2, N=4200, S=256, H=160, W=192)),
])
def test_forward_matches_reference_golden(name, c):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    model = _build(c["frames"])
    sample = orc.make_inputs(seed=1, B=1, T=c["T"], N=c["N"], S=c["S"], H=c["H"], W=c["W"])
    ret = _run(model, sample)
    assert isinstance(ret, dict) and "pcd_moved" in ret and tuple(ret["pcd_moved"].shape) == (1, c["T"], c["N"], 3)
    ref = torch.from_numpy(g["pcd_moved"])
    err = orc.rel_l2(ret.pcd_moved.cpu(), ref)
    assert err < REL_TOL, f"{name}: rel-L2 {err:.3e} vs reference golden"
    assert abs(float(ret.loss_metrics.loss) - float(g["loss"])) < REL_TOL * abs(float(g["loss"]))
    assert abs(float(ret.loss_metrics.xyz_loss) - float(g["xyz_loss"])) < REL_TOL * abs(float(g["xyz_loss"]))


def test_forward_config_b_matches_reference_golden():
    """BASELINE config (b), the workload bench.py times (32 frames x 4096 points, S = 4096, weights seed 0, inputs seed 1):
    pcd_moved and loss against the UNMODIFIED reference's fp32 forward (tests/golden/make_golden.py b_T32_N4096: every 4th
    point of every frame stored + fp64 checksums over all points).  bench.py asserts its printed loss against the same run."""
    g = np.load(os.path.join(GOLD, "b_T32_N4096.npz"))
    T, N, S = 32, 4096, 4096
    model = _build(T)
    ret = _run(model, orc.make_inputs(seed=1, B=1, T=T, N=N, S=S))
    assert tuple(ret.pcd_moved.shape) == (1, T, N, 3)
    stride = int(g["point_stride"])
    got, ref = ret.pcd_moved.cpu(), torch.from_numpy(g["pcd_moved"])
    err = orc.rel_l2(got[:, :, ::stride], ref)
    worst_frame = max(orc.rel_l2(got[:, t, ::stride], ref[:, t]) for t in range(T))
    print(f"config (b): pcd_moved rel-L2 {err:.3e} (worst frame {worst_frame:.3e}), loss {float(ret.loss_metrics.loss):.8f} vs {float(g['loss']):.8f}")
    assert err < REL_TOL and worst_frame < 2 * REL_TOL, (err, worst_frame)
    # checksums over ALL points (the stored subset is every 4th): sum and sum of squares
    assert abs(float(got.double().sum()) - float(g["pcd_moved_sum"])) < REL_TOL * float(g["pcd_moved_sqsum"]) ** 0.5 * (got.numel() ** 0.5)
    assert abs(float(got.double().pow(2).sum()) - float(g["pcd_moved_sqsum"])) < 2 * REL_TOL * float(g["pcd_moved_sqsum"])
    lref = float(g["loss"])
    assert abs(float(ret.loss_metrics.loss) - lref) < REL_TOL * lref
    import bench
    assert abs(bench.LOSS_FP32_ORACLE - lref) < 1e-6 * lref     # the constant bench.py checks its loss against is this run's


def test_forward_stagewise_vs_oracle_T4_batch2():
    """B=2, T=4: every stage boundary against the oracle; also exercises batch > 1 and frame-chunked decoding."""
    frames, T, N, S = 4, 4, 640, 512
    model = _build(frames)
    model.max_decode_rows = 2 * N  # force 2 decoder chunks per batch element
    sample = orc.make_inputs(seed=3, B=2, T=T, N=N, S=S)
    ret = _run(model, sample)
    with torch.no_grad():
        ref = orc.forward(orc.init_state_dict(0, dict(frames=frames)), sample, dict(frames=frames), return_stages=True)
    st = ref["stages"]
    ws = {k[0]: v for k, v in model._ws.items()}
    assert orc.rel_l2(ws["mesh_feat"].cpu().view(2, 64, 768), st["mesh_feat"]) < REL_TOL
    assert orc.rel_l2(ws["trunk_x"].cpu().view(2, T, 324, 768), st["trunk_out"]) < REL_TOL
    err = orc.rel_l2(ret.pcd_moved.cpu(), ref["pcd_moved"])
    assert err < REL_TOL, f"rel-L2 {err:.3e}"
    assert abs(float(ret.loss_metrics.loss) - float(ref["loss_metrics"]["loss"])) < REL_TOL * float(ref["loss_metrics"]["loss"])
    # deterministic: same inputs -> bit-identical outputs
    ret2 = _run(model, sample)
    assert torch.equal(ret.pcd_moved, ret2.pcd_moved) and torch.equal(ret.loss_metrics.loss, ret2.loss_metrics.loss)


def test_no_ground_truth_and_errors():
    model = _build(1)
    sample = orc.make_inputs(seed=1, B=1, T=1, N=256, S=256, with_gt=False)
    ret = _run(model, sample)
    assert "loss_metrics" not in ret and tuple(ret.pcd_moved.shape) == (1, 1, 256, 3)
    bad = {k: v.to("cuda") for k, v in orc.make_inputs(seed=1, B=1, T=1, N=256, S=256).items()}
    bad["point_clouds"] = bad["point_clouds"][:, :, :100]
    with pytest.raises(ValueError):
        model(bad)
    with pytest.raises(RuntimeError):
        model(orc.make_inputs(seed=1, B=1, T=1, N=16, S=16))  # CPU tensors: no CPU path


def test_smooth_trajectories_matches_reference_golden():
    """GPU smoothing vs the output of the reference's own smooth_trajectories(method='combined')."""
    from motion324_b200.inference import smooth_trajectories
    g = np.load(os.path.join(GOLD, "inference_smooth.npz"))
    trajs = torch.from_numpy(g["trajs"]).to("cuda")
    out = smooth_trajectories(trajs, method="combined", motion_threshold=0.002, sigma=1.0)
    ref = torch.from_numpy(g["smoothed"])
    assert float((out.cpu() - ref).abs().max()) < 2e-7          # fp64 accumulation, fp32 result: at most one rounding apart
    thr = smooth_trajectories(trajs, method="threshold", motion_threshold=0.002)
    from oracle import inference_oracle as io
    assert torch.equal(thr.cpu(), io.smooth_trajectories(torch.from_numpy(g["trajs"]), 0.002, 1.0, method="threshold"))  # pure selection: exact
    big = torch.randn(1, 64, 20000, 3, device="cuda").cumsum(1) * 0.003
    o = smooth_trajectories(big, method="combined", motion_threshold=0.002, sigma=1.0)
    assert orc.rel_l2(o.cpu(), io.smooth_trajectories(big.cpu(), 0.002, 1.0)) < 1e-6
    # the other methods of the same function, against the reference function's own outputs
    small = torch.from_numpy(g["small"]).to("cuda")
    for w, po in ((3, 2), (5, 2), (7, 3), (4, 2)):
        got = smooth_trajectories(small, method="savgol", window_size=w, savgol_polyorder=po)
        assert float((got.cpu() - torch.from_numpy(g[f"savgol_w{w}_p{po}"])).abs().max()) < 2e-7, (w, po)      # fp64 sums, one fp32 rounding
    assert torch.equal(smooth_trajectories(small, method="oneeuro").cpu(), torch.from_numpy(g["oneeuro_default"]))           # fp32 recurrence: bit-exact
    assert torch.equal(smooth_trajectories(small, method="oneeuro", oneeuro_mincutoff=0.3, oneeuro_beta=0.5).cpu(), torch.from_numpy(g["oneeuro_b05"]))
    assert float((smooth_trajectories(small, method="gaussian", sigma=1.5).cpu() - torch.from_numpy(g["gaussian_s15"])).abs().max()) < 2e-7
    assert torch.equal(smooth_trajectories(small, method="none-of-them"), small)                 # unknown method: a copy, like the reference
    short = small[:, :2].contiguous()
    assert torch.equal(smooth_trajectories(short, method="savgol", window_size=5), short)        # T < window: untouched (:153)


def test_run_model_inference_windows_and_stitch():
    """Sliding-window inference (inference_with_video_mesh.py:132-256) through the real model vs per-window oracle stitching."""
    from motion324_b200.inference import run_model_inference, window_plan
    from oracle import inference_oracle as io
    frames, total_T, N, S = 3, 7, 200, 256
    model = _build(frames)
    sample = orc.make_inputs(seed=5, B=1, T=total_T, N=N, S=S, with_gt=False)
    video = sample.pop("rgb_video")[0]
    inp = {k: v.to("cuda") for k, v in sample.items()}
    trajs = run_model_inference(model, inp, video, model.config, "cuda")
    plan = window_plan(total_T, frames)
    assert [f for _, f in plan] == [f for _, f in io.window_plan(total_T, frames)]
    assert tuple(trajs.shape) == (1, total_T, N, 3)
    assert torch.equal(trajs[:, 0], inp["ref_pcd"])             # frame 0 is the reference shape, bit-exact
    sd = orc.init_state_dict(0, dict(frames=frames))
    outs = []
    with torch.no_grad():
        for _, fr in plan:
            s = dict(sample)
            s["rgb_video"] = video[fr][None]
            outs.append(orc.forward(sd, s, dict(frames=frames))["pcd_moved"])
    ref = io.stitch(outs, [s for s, _ in plan], sample["ref_pcd"])
    assert orc.rel_l2(trajs.cpu(), ref) < REL_TOL


def test_cuda_graph_replay_is_bit_identical_to_the_eager_forward():
    """include/m324.h promises capture-safe launches: the whole inference forward (both streams, ~270 launches) captured into
    ONE CUDA graph replays bit-identically to the eager launch sequence, on new inputs of the same shapes, with zero library
    launches issued from the host during a replay; a second shape signature gets its own graph; a weight update drops them."""
    from motion324_b200 import ops
    frames, T, N, S = 3, 3, 300, 280
    model = _build(frames)
    s1 = {k: v.to("cuda") for k, v in orc.make_inputs(seed=11, B=1, T=T, N=N, S=S).items()}
    s2 = {k: v.to("cuda") for k, v in orc.make_inputs(seed=12, B=1, T=T, N=N, S=S).items()}
    e1, e2 = model(s1), model(s2)
    e1 = (e1.pcd_moved.clone(), e1.loss_metrics.loss.clone())
    e2 = (e2.pcd_moved.clone(), e2.loss_metrics.loss.clone())
    model.enable_cuda_graph(True)
    g1 = model(s1)                       # captures
    assert torch.equal(g1.pcd_moved, e1[0]) and torch.equal(g1.loss_metrics.loss, e1[1])
    n0 = ops.launch_count()
    g2 = model(s2)                       # replays on new inputs
    torch.cuda.synchronize()
    assert ops.launch_count() == n0      # nothing launched from the host
    assert torch.equal(g2.pcd_moved, e2[0]) and torch.equal(g2.loss_metrics.loss, e2[1])
    assert len(model._graphs) == 1
    s3 = {k: v.to("cuda") for k, v in orc.make_inputs(seed=13, B=1, T=T, N=200, S=S).items()}
    g3 = model(s3)
    assert len(model._graphs) == 2 and tuple(g3.pcd_moved.shape) == (1, T, 200, 3)
    model.enable_cuda_graph(False)
    e3 = model(s3)
    assert torch.equal(g3.pcd_moved, e3.pcd_moved)
    model.enable_cuda_graph(True)
    model(s1)
    with torch.no_grad():
        model.shared_mlp_output[3].bias.add_(1.0)        # in-place weight update (what optimizer.step() does)
    g4 = model(s1)
    assert torch.allclose(g4.pcd_moved, e1[0] + 1.0, atol=1e-5)


def test_back_to_back_forwards_are_bit_identical_at_config_b():
    """The benchmark loop: forwards issued back to back with no host sync in between (kernels of consecutive steps overlap through PDL,
    the persistent attention / GEMM kernels see different ring timings) must return the bits of a forward that ran alone.  Guards the
    pipeline protocols: a barrier-phase aliasing in the item-loop attention kernel once showed up ONLY here (scripts/determinism_check.py)."""
    T, N = 32, 4096
    model = _build(T)
    sample = {k: v.to("cuda") for k, v in orc.make_inputs(seed=1, B=1, T=T, N=N, S=N).items()}
    r = model(sample)
    torch.cuda.synchronize()
    ref, ref_loss = r.pcd_moved.clone(), r.loss_metrics.loss.clone()
    for burst in range(3):
        outs = []
        for _ in range(10):
            r = model(sample)
            outs.append((r.pcd_moved.clone(), r.loss_metrics.loss.clone()))
        torch.cuda.synchronize()
        bad = [i for i, (o, l) in enumerate(outs) if not (torch.equal(o, ref) and torch.equal(l, ref_loss))]
        assert not bad, f"burst {burst}: forwards {bad} differ from the forward that ran alone"
