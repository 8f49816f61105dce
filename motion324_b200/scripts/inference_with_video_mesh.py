"""Entry point mirroring /root/reference/scripts/inference_with_video_mesh.py (animate an existing mesh with a video) for the
steps either side of the hot path, same function names and argument meaning:

    load_video_from_path(video_path)                  (:26-57)    cv2.VideoCapture instead of imageio (absent here)
    prepare_mesh_data(config, glb_path, device)       (:60-129)   own GLB reader; surface sampling on the host, texture lookup
                                                                   (m324_sample_albedo) and nearest-sample colour transfer
                                                                   (m324_chamfer_nn instead of scipy cKDTree) on the GPU
    run_model_inference(model, input_data, video, config, device)  (:132-256)  motion324_b200.inference
    run_inference_on_video(config)                    (:301-430)  up to the smoothed trajectories; the Blender (bpy) shape-key
                                                                   export, rembg segmentation and the FBX conversion are OUT OF
                                                                   SCOPE: use_segmentation must be False and
                                                                   the trajectories are written to <output_dir>/trajs.npy.

    python -m motion324_b200.scripts.inference_with_video_mesh --config configs/dyscene.yaml data_dir=a.glb video_path=a.mp4 ...
"""
import importlib
import os

import numpy as np
import torch

from ..inference import run_model_inference, smooth_trajectories   # noqa: F401  (same names as the reference module)
from ..utils.mesh_processing import SimpleMesh, sample_pointcloud_with_albedo


def load_video_from_path(video_path):
    """-> uint8 [T, H, W, 3] RGB in [0, 255]."""
    import cv2
    if video_path.lower().endswith((".mp4", ".avi", ".mov")):
        cap = cv2.VideoCapture(video_path)
        if not cap.isOpened():
            raise ValueError(f"cannot open video: {video_path}")
        frames = []
        while True:
            ok, frame = cap.read()
            if not ok:
                break
            frames.append(frame[:, :, ::-1])
        cap.release()
    elif os.path.isdir(video_path):
        names = sorted((f for f in os.listdir(video_path) if f.lower().endswith((".png", ".jpg", ".jpeg"))),
                       key=lambda s: [int(t) if t.isdigit() else t for t in __import__("re").split(r"(\d+)", s)])   # natural order
        frames = [cv2.imread(os.path.join(video_path, f), cv2.IMREAD_COLOR)[:, :, ::-1] for f in names]
    else:
        raise ValueError(f"video_path must be a video file or image directory: {video_path}")
    if not frames:
        raise ValueError(f"no frames in {video_path}")
    return np.ascontiguousarray(np.stack(frames, axis=0))


def nearest_sample_indices(samples_xyz, vertices, device):
    """cKDTree(samples).query(vertices, k=1)[1] (:113-115) as an exact brute-force search on the GPU (ties -> smallest index)."""
    import ctypes as C
    from .. import lib as _l, ops as _ops
    dev = torch.device(device)
    s = torch.as_tensor(samples_xyz, dtype=torch.float64).to(dev).contiguous()
    v = torch.as_tensor(vertices, dtype=torch.float64).to(dev).contiguous()
    n1, n2 = s.shape[0], v.shape[0]
    d1 = torch.empty(n2, device=dev, dtype=torch.float64)
    i1 = torch.empty(n2, device=dev, dtype=torch.int32)
    d2 = torch.empty(n1, device=dev, dtype=torch.float64)
    with torch.cuda.device(dev):
        _ops.LAUNCHES[0] += 1
        _l.check(_l.load().m324_chamfer_nn(C.c_void_p(s.data_ptr()), n1, C.c_void_p(v.data_ptr()), n2, 1, 1, C.c_void_p(d1.data_ptr()),
                                           C.c_void_p(i1.data_ptr()), C.c_void_p(d2.data_ptr()), None, _ops._stream()), "m324_chamfer_nn")
    return i1.long()


def prepare_mesh_data(config, glb_path, device, rng=777):
    """-> (input_data dict of [1, ., 3] float tensors on `device` + 'faces', mesh, faces) like the reference."""
    mesh = SimpleMesh.from_glb(glb_path)
    vertices = mesh.vertices.astype(np.float32)
    faces = mesh.faces.astype(np.int64)
    vertex_normals_np = mesh.vertex_normals.astype(np.float32)
    center = (vertices.max(axis=0) + vertices.min(axis=0)) / 2          # :94-97
    vertices = vertices - center
    v_max = np.abs(vertices).max()
    vertices = vertices / (2 * (v_max + 1e-8))
    mesh.vertices = (mesh.vertices - center) / (2 * (v_max + 1e-8))     # :103-105 (float64 like trimesh's array)
    tr = config.training
    num = tr.get("num_shape_samples", 16384) if hasattr(tr, "get") else getattr(tr, "num_shape_samples", 16384)
    samples_xyz, samples_normals, samples_rgb = sample_pointcloud_with_albedo(mesh, num=num, rng=rng, device=device)
    nearest = nearest_sample_indices(samples_xyz.numpy(), vertices, device)      # :113-116
    vert_rgb = samples_rgb.to(device)[nearest]
    input_data = {
        "ref_shape_pcd": samples_xyz[None].float().to(device),
        "ref_shape_normals": samples_normals[None].float().to(device),
        "ref_shape_rgbs": samples_rgb[None].float().to(device),
        "ref_pcd": torch.from_numpy(vertices)[None].float().to(device),
        "ref_normal": torch.from_numpy(vertex_normals_np)[None].float().to(device),
        "ref_rgb": vert_rgb[None].float(),
        "faces": torch.from_numpy(faces)[None].long().to(device),
    }
    return input_data, mesh, faces


def select_frames(video_np, T, start_frame=0):
    """:371-383: float [T', H, W, 3] in [0, 1], the first T frames from start_frame (fewer when the clip is shorter)."""
    video = torch.from_numpy(video_np.astype(np.float32)).float() / 255.0
    return video[start_frame:start_frame + T] if start_frame + T <= video.shape[0] else video[start_frame:]


def run_inference_on_video(config, model=None):
    """:301-407 without segmentation and without the Blender export.  Returns the smoothed trajectories [1, T, V, 3]."""
    device = torch.device("cuda", torch.cuda.current_device())
    if getattr(config, "use_segmentation", True):
        raise NotImplementedError("rembg / U2Net foreground segmentation is out of scope: pass use_segmentation=False")
    if model is None:
        module, class_name = config.model.class_name.rsplit(".", 1)
        model = importlib.import_module(module).__dict__[class_name](config).to(device)
        ckpt = config.training.get("resume_ckpt", "")
        if ckpt:
            sd = torch.load(ckpt, map_location=device)["model"]
            sd.pop("pos_embed", None)                                   # utils/inference_utils.py:39-40
            try:
                model.load_state_dict(sd, strict=True)
            except RuntimeError:
                model.load_state_dict(sd, strict=False)
    model.eval()
    with torch.no_grad():
        input_data, mesh, faces = prepare_mesh_data(config, config.data_dir, device)
        video = select_frames(load_video_from_path(config.video_path), config.training.frames, getattr(config, "start_frame", 0))
        input_data["rgb_video"] = video
        trajs = run_model_inference(model, input_data, video, config, device)
        if trajs is not None:
            trajs = smooth_trajectories(trajs, method="combined", motion_threshold=0.002, window_size=3, sigma=1.0)
    out_dir = getattr(config, "output_dir", None)
    if out_dir and trajs is not None:
        os.makedirs(out_dir, exist_ok=True)
        np.save(os.path.join(out_dir, "trajs.npy"), trajs.cpu().numpy())
    return trajs


if __name__ == "__main__":
    from ..utils.config import init_config
    run_inference_on_video(init_config())
