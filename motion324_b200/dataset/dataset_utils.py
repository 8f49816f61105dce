"""GPU mirror of the reference's data-prep gathers (/root/reference/dataset/dataset_utils.py).

Same names and argument meaning as the reference module for the parts that are arithmetic:

* ``sample_texture_color_vectorized(uvs, texture_array)``           (dataset_utils.py:19-41)
* ``track_with_normal_rgb(...)``                                     (dataset_utils.py:44-136) -- the reference samples the
  surface with trimesh's RNG inside the function; here the sampled ``face_indices`` / ``barycentric_coords`` (and trimesh's
  per-frame ``vertex_normals``) are ARGUMENTS, because surface sampling and vertex-normal estimation are trimesh code, not
  Motion324 code, and stay on the host.  Everything after them -- the index gathers, barycentric interpolation through all
  frames, normal interpolation + normalisation, UV interpolation and the texture lookup -- runs in libm324 (m324_track_points,
  m324_sample_texture_colors) in float64 like NumPy, rounded to float32 once at the end like the reference (:131-133).

CUDA only: there is no CPU path (a CPU tensor raises)."""
import ctypes as C

import numpy as np
import torch

from .. import lib as _l
from .. import ops as _ops


def _dev(x, dtype, device):
    t = torch.as_tensor(np.ascontiguousarray(x) if isinstance(x, np.ndarray) else x)
    return t.to(device=device, dtype=dtype).contiguous()


def _check(err, what):
    code = int(err.item())
    if code:
        raise IndexError(f"{what}: {'face' if code == 1 else 'vertex'} index out of range")   # NumPy's fancy indexing raises too


def sample_texture_color_vectorized(uvs, texture_array, device="cuda"):
    """Texture lookup at UV coordinates (dataset_utils.py:19-41): returns the [N, 3] uint8 texels ``texture_array[y, x]``.
    Implemented as the degenerate case of m324_sample_texture_colors (one 'face' per sample with all three corners = uv)."""
    uvs = _dev(uvs, torch.float64, device)
    n = uvs.shape[0]
    tex = _dev(texture_array, torch.uint8, device)
    assert tex.ndim == 3 and tex.shape[2] == 3 and uvs.shape == (n, 2)
    face_uvs = uvs[:, None, :].expand(n, 3, 2).contiguous()
    bary = torch.zeros(n, 3, device=uvs.device, dtype=torch.float64)
    bary[:, 0] = 1.0
    rgb, texel = _sample(face_uvs, torch.arange(n, device=uvs.device), bary, tex)
    return tex[texel[:, 0], texel[:, 1]] if n else tex.new_zeros(0, 3)


def _sample(face_uvs, face_indices, bary, tex):
    dev = face_uvs.device
    if not face_uvs.is_cuda:
        raise RuntimeError("motion324_b200 data prep runs on CUDA tensors only: there is no CPU path")
    n = face_indices.shape[0]
    rgb = torch.empty(n, 3, device=dev, dtype=torch.float32)
    texel = torch.empty(n, 2, device=dev, dtype=torch.int64)
    err = torch.zeros(1, device=dev, dtype=torch.int32)
    _ops.LAUNCHES[0] += 1
    _l.check(_l.load().m324_sample_texture_colors(C.c_void_p(face_uvs.data_ptr()), face_uvs.shape[0], C.c_void_p(face_indices.data_ptr()),
                                                  C.c_void_p(bary.data_ptr()), n, C.c_void_p(tex.data_ptr()), tex.shape[0], tex.shape[1],
                                                  C.c_void_p(rgb.data_ptr()), C.c_void_p(texel.data_ptr()), C.c_void_p(err.data_ptr()),
                                                  _ops._stream()), "m324_sample_texture_colors")
    _check(err, "sample_texture_colors")
    return rgb, texel


def track_with_normal_rgb(vertex_frames, faces, face_indices, barycentric_coords, face_uvs, texture_array, vertex_normals=None,
                          device="cuda"):
    """dataset_utils.py:44-136 after the trimesh sampling.  vertex_frames [T, V, 3]; faces [F, 3]; face_indices [S];
    barycentric_coords [S, 3]; face_uvs [F, 3, 2]; texture_array [H, W, 3] uint8; vertex_normals [T, V, 3] (trimesh's
    ``vertex_normals`` per frame) or None.  Returns (tracked_points [T, S, 3], tracked_normals [T, S, 3] or None,
    tracked_rgbs [T, S, 3], face_indices) as float32 CUDA tensors."""
    f64 = (vertex_frames.dtype == np.float64) if isinstance(vertex_frames, np.ndarray) else (vertex_frames.dtype == torch.float64)
    vdt = torch.float64 if f64 else torch.float32
    verts = _dev(vertex_frames, vdt, device)
    if not verts.is_cuda:
        raise RuntimeError("motion324_b200 data prep runs on CUDA tensors only: there is no CPU path")
    T, V = verts.shape[:2]
    vn = _dev(vertex_normals, vdt, device) if vertex_normals is not None else None
    faces_d = _dev(faces, torch.int64, device)
    fidx = _dev(face_indices, torch.int64, device)
    bary = _dev(barycentric_coords, torch.float64, device)
    S = fidx.shape[0]
    points = torch.empty(T, S, 3, device=verts.device, dtype=torch.float32)
    normals = torch.empty(T, S, 3, device=verts.device, dtype=torch.float32) if vn is not None else None
    err = torch.zeros(1, device=verts.device, dtype=torch.int32)
    _ops.LAUNCHES[0] += 1
    _l.check(_l.load().m324_track_points(C.c_void_p(verts.data_ptr()), C.c_void_p(vn.data_ptr()) if vn is not None else None, int(f64), T, V,
                                         C.c_void_p(faces_d.data_ptr()), faces_d.shape[0], C.c_void_p(fidx.data_ptr()),
                                         C.c_void_p(bary.data_ptr()), S, C.c_void_p(points.data_ptr()),
                                         C.c_void_p(normals.data_ptr()) if normals is not None else None, C.c_void_p(err.data_ptr()),
                                         _ops._stream()), "m324_track_points")
    _check(err, "track_points")
    rgb, _ = _sample(_dev(face_uvs, torch.float64, device), fidx, bary, _dev(texture_array, torch.uint8, device))
    return points, normals, rgb[None].expand(T, S, 3), face_indices     # colours are fixed across frames (:129-133)
