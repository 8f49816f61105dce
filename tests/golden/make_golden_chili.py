"""Golden fixture for BASELINE config (d): inference on examples/chili.glb + chili.mp4 (T = 159 frames, 13,465 vertices,
S = 16,384 surface samples, ``training.frames = 256`` -> the trilinear pos-embed resize path, 4 decoder chunks).

Runs only in the build container.  What is the reference's own code here and what is not:
  * surface-sample colours: the UNMODIFIED reference functions ``barycentric_coords`` / ``sample_pointcloud_with_albedo``
    (utils/mesh_processing.py:107-191, extracted with ``ast`` because the module imports trimesh) on a stand-in mesh object that
    replays the sampled points (trimesh itself is absent: GLB parsing and the surface sampler are ours, parity unpinned there);
  * vertex colours: scipy cKDTree nearest sample, as scripts/inference_with_video_mesh.py:113-116;
  * trajectories: the UNMODIFIED reference ``Motion_Latent_Model`` (fp32, CPU, exact attention) on those tensors + the video
    frames decoded with OpenCV (imageio is absent), weights = motion324_b200.utils.synthetic.init_state_dict(0) with the
    ``pos_embed`` of a 256-frame model -- one model() call, as run_model_inference does for total_T <= chunk (:157-174).
Stored: the sample colours as bytes (exact), the nearest-sample indices, every 32nd vertex of pcd_moved + fp64 checksums.

    python tests/golden/make_golden_chili.py
"""
import ast
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import build_ref, ref_shims  # noqa: E402
from motion324_b200.utils import synthetic as syn  # noqa: E402
from motion324_b200.utils.mesh_processing import SimpleMesh  # noqa: E402
from motion324_b200.scripts.inference_with_video_mesh import load_video_from_path, select_frames  # noqa: E402

FRAMES_CFG, S_SAMPLES, SEED, VSTRIDE = 256, 16384, 777, 32


def reference_functions():
    src = open(os.path.join(build_ref.root(), "utils", "mesh_processing.py")).read()
    tree = ast.parse(src)
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("barycentric_coords", "sample_pointcloud_with_albedo")]
    from PIL import Image
    ns = {"np": np, "torch": torch, "Image": Image}
    exec(compile(ast.Module(body=keep, type_ignores=[]), "mesh_processing.py", "exec"), ns)
    return ns


class _Visual:
    pass


class ReplayMesh:
    """The attributes sample_pointcloud_with_albedo reads from a trimesh.Trimesh, with ``sample`` replaying fixed samples."""

    def __init__(self, mesh, points, face_idx):
        from PIL import Image
        self.vertices, self.faces, self.triangles, self.face_normals = mesh.vertices, mesh.faces, mesh.triangles, mesh.face_normals
        self._s = (points, face_idx)
        self.visual = _Visual()
        self.visual.uv = mesh.uv
        self.visual.material = _Visual()
        self.visual.material.baseColorTexture = Image.fromarray(mesh.texture)

    def sample(self, num, return_index=False):
        assert num == len(self._s[0]) and return_index
        return self._s


def build_inputs():
    """The tensors prepare_mesh_data produces, computed with the reference's own arithmetic on the host."""
    from scipy.spatial import cKDTree
    root = build_ref.root()
    mesh = SimpleMesh.from_glb(os.path.join(root, "examples", "chili.glb"))
    vertices = mesh.vertices.astype(np.float32)
    normals = mesh.vertex_normals.astype(np.float32)
    center = (vertices.max(axis=0) + vertices.min(axis=0)) / 2
    vertices = vertices - center
    v_max = np.abs(vertices).max()
    vertices = vertices / (2 * (v_max + 1e-8))
    mesh.vertices = (mesh.vertices - center) / (2 * (v_max + 1e-8))
    points, face_idx = mesh.sample(S_SAMPLES, return_index=True, rng=SEED)
    ns = reference_functions()
    xyz, nrm, rgb = ns["sample_pointcloud_with_albedo"](ReplayMesh(mesh, points, face_idx), num=S_SAMPLES)
    _, nearest = cKDTree(xyz.numpy()).query(vertices, k=1)
    vert_rgb = rgb[nearest]
    video = select_frames(load_video_from_path(os.path.join(root, "examples", "chili.mp4")), FRAMES_CFG)
    sample = dict(ref_shape_pcd=xyz[None].float(), ref_shape_normals=nrm[None].float(), ref_shape_rgbs=rgb[None].float(),
                  ref_pcd=torch.from_numpy(vertices)[None].float(), ref_normal=torch.from_numpy(normals)[None].float(),
                  ref_rgb=vert_rgb[None].float(), rgb_video=video[None].float())
    return sample, face_idx, nearest


def main():
    torch.set_num_threads(os.cpu_count())
    sample, face_idx, nearest = build_inputs()
    T, N = sample["rgb_video"].shape[1], sample["ref_pcd"].shape[1]
    print("inputs:", {k: tuple(v.shape) for k, v in sample.items()})
    model = ref_shims.build_reference_model(frames=FRAMES_CFG)
    model.load_state_dict(syn.init_state_dict(0, dict(frames=FRAMES_CFG)), strict=True)
    t0 = time.time()
    with torch.no_grad():
        ret = model(dict(sample))
    out = ret["pcd_moved"].float()
    print(f"reference forward: {time.time() - t0:.0f} s, pcd_moved {tuple(out.shape)}")
    rgb8 = np.rint(sample["ref_shape_rgbs"][0].numpy() * 255.0).astype(np.uint8)
    assert np.array_equal(rgb8.astype(np.float32) / np.float32(255.0), sample["ref_shape_rgbs"][0].numpy())     # bytes are exact
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "chili_T159.npz"),
                        face_idx=face_idx.astype(np.int32), shape_rgb_u8=rgb8, nearest=nearest.astype(np.int32),
                        shape_pcd_sum=np.float64(sample["ref_shape_pcd"].double().sum().item()),
                        video_sum=np.float64(sample["rgb_video"].double().sum().item()),
                        pcd_moved=np.ascontiguousarray(out[:, :, ::VSTRIDE].numpy()), vertex_stride=np.int64(VSTRIDE),
                        pcd_moved_sum=np.float64(out.double().sum().item()), pcd_moved_sqsum=np.float64(out.double().pow(2).sum().item()),
                        T=np.int64(T), N=np.int64(N))
    print("saved chili_T159.npz")


if __name__ == "__main__":
    main()
