"""Mesh -> tensors step in front of the hot path at inference (second half): the parts of
/root/reference/utils/mesh_processing.py that scripts/inference_with_video_mesh.py:60-129 uses, same names and argument meaning.

* ``SimpleMesh``                      -- the handful of ``trimesh.Trimesh`` attributes those call sites read (vertices, faces,
                                         vertex_normals, face_normals, triangles, area_faces, uv, texture, ``sample``), built from
                                         ``motion324_b200.utils.glb.load_glb``.  trimesh is absent in this image; its surface sampler
                                         is restated from its published algorithm (area-weighted face pick, folded random
                                         barycentrics) with an explicit NumPy Generator -- trimesh internals are parity-UNPINNED.
* ``sample_pointcloud_with_albedo``   -- mesh_processing.py:130-191; the per-sample Python loop (:174-182) runs in libm324
                                         (``m324_sample_albedo``), texel indices bit-exact against the reference loop.
* ``normalize_mesh``                  -- mesh_processing.py:194-218 (host NumPy; a min / max and a divide).
"""
import ctypes as C

import numpy as np
import torch

from .. import lib as _l
from .. import ops as _ops
from . import glb as _glb


class SimpleMesh:
    def __init__(self, vertices, faces, vertex_normals=None, uv=None, texture=None):
        self.vertices = np.ascontiguousarray(vertices, dtype=np.float64)
        self.faces = np.ascontiguousarray(faces, dtype=np.int64)
        self._vertex_normals = None if vertex_normals is None else np.ascontiguousarray(vertex_normals, dtype=np.float64)
        self.uv = None if uv is None else np.ascontiguousarray(uv, dtype=np.float64)
        self.texture = texture          # [H, W, 3] uint8 RGB or None

    @classmethod
    def from_glb(cls, path):
        g = _glb.load_glb(path)
        return cls(g["vertices"], g["faces"], g["normals"], g["uv"], g["texture"])

    @property
    def triangles(self):
        return self.vertices[self.faces]

    def _cross(self):
        t = self.triangles
        return np.cross(t[:, 1] - t[:, 0], t[:, 2] - t[:, 0])

    @property
    def area_faces(self):
        return np.linalg.norm(self._cross(), axis=1) / 2.0

    @property
    def face_normals(self):
        c = self._cross()
        n = np.linalg.norm(c, axis=1, keepdims=True)
        return c / np.where(n == 0.0, 1.0, n)

    @property
    def vertex_normals(self):
        if self._vertex_normals is not None:
            return self._vertex_normals
        acc = np.zeros_like(self.vertices)          # area-weighted face normals, the common fallback
        c = self._cross()
        for k in range(3):
            np.add.at(acc, self.faces[:, k], c)
        n = np.linalg.norm(acc, axis=1, keepdims=True)
        return acc / np.where(n == 0.0, 1.0, n)

    def sample(self, count, return_index=False, rng=None):
        """Uniform surface samples (trimesh.sample.sample_surface): faces picked with probability proportional to area, a
        uniform point in the chosen triangle by folding two uniform edge lengths.  rng: numpy Generator or seed."""
        rng = np.random.default_rng(rng)
        cum = np.cumsum(self.area_faces)
        face_index = np.searchsorted(cum, rng.random(count) * cum[-1])
        tri = self.triangles[face_index]
        origin, vec = tri[:, 0], tri[:, 1:] - tri[:, :1]
        lengths = rng.random((count, 2, 1))
        fold = lengths.sum(axis=1).reshape(-1) > 1.0
        lengths[fold] -= 1.0
        lengths = np.abs(lengths)
        points = (vec * lengths).sum(axis=1) + origin
        return (points, face_index) if return_index else points


def normalize_mesh(mesh, return_params=False):
    """mesh_processing.py:194-218."""
    vertices = mesh.vertices.astype(np.float32)
    center = (vertices.max(axis=0) + vertices.min(axis=0)) / 2
    vertices = vertices - center
    v_max = np.abs(vertices).max()
    scale = 2 * (v_max + 1e-8)
    vertices = vertices / scale
    if return_params:
        return vertices, center, scale
    mesh.vertices = np.ascontiguousarray(vertices, dtype=np.float64)
    return mesh


def sample_albedo(mesh, points, face_idx, device="cuda", return_texels=False):
    """The loop of mesh_processing.py:174-182 on the GPU: colours [S, 3] float32 (and the gathered (y, x) texel indices)."""
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("motion324_b200 mesh processing runs on CUDA only: there is no CPU path")
    f64 = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
    verts, uv, pts = f64(mesh.vertices), f64(mesh.uv), f64(points)
    faces = torch.as_tensor(mesh.faces).to(dev)
    fidx = torch.as_tensor(np.ascontiguousarray(face_idx, dtype=np.int64)).to(dev)
    tex = torch.as_tensor(np.ascontiguousarray(mesh.texture)).to(dev)
    S = fidx.shape[0]
    rgb = torch.empty(S, 3, device=dev, dtype=torch.float32)
    texel = torch.empty(S, 2, device=dev, dtype=torch.int64)
    err = torch.zeros(1, device=dev, dtype=torch.int32)
    with torch.cuda.device(dev):
        _ops.LAUNCHES[0] += 1
        _l.check(_l.load().m324_sample_albedo(C.c_void_p(verts.data_ptr()), verts.shape[0], C.c_void_p(faces.data_ptr()), faces.shape[0],
                                              C.c_void_p(uv.data_ptr()), C.c_void_p(fidx.data_ptr()), C.c_void_p(pts.data_ptr()), S,
                                              C.c_void_p(tex.data_ptr()), tex.shape[0], tex.shape[1], C.c_void_p(rgb.data_ptr()),
                                              C.c_void_p(texel.data_ptr()), C.c_void_p(err.data_ptr()), _ops._stream()), "m324_sample_albedo")
    code = int(err.item())
    if code:
        raise IndexError(f"sample_albedo: {'face' if code == 1 else 'vertex'} index out of range")
    return (rgb, texel) if return_texels else rgb


def sample_pointcloud_with_albedo(mesh, num=200000, rng=None, device="cuda"):
    """mesh_processing.py:130-191: (points [num, 3], face normals [num, 3], colours [num, 3] in [0, 1]) as float32 tensors
    (points / normals on the host like the reference, colours from the GPU kernel moved back)."""
    points, face_idx = mesh.sample(num, return_index=True, rng=rng)
    normals = mesh.face_normals[face_idx]
    if mesh.texture is not None and mesh.uv is not None:
        colors = sample_albedo(mesh, points, face_idx, device).cpu()
    else:
        colors = torch.full((num, 3), 0.5, dtype=torch.float32)        # :184-185
    return torch.from_numpy(points.astype(np.float32)), torch.from_numpy(normals.astype(np.float32)), colors
