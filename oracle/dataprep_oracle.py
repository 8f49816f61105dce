"""ORACLE (test infrastructure, never on the product path): NumPy restatement of the reference's data-prep arithmetic,
/root/reference/dataset/dataset_utils.py, for SURVEY.md 8(f4).

Pinned by tests/golden/dataprep.npz, which tests/golden/make_golden_dataprep.py produced by running the UNMODIFIED reference
functions (``sample_texture_color_vectorized`` as is; ``track_with_normal_rgb`` with a stand-in for the absent ``trimesh``
package that replays fixed samples).  PARITY UNPINNED for what lives inside trimesh (un-vendored dependency, requirements
list it without a version): ``sample_surface``'s RNG stream and ``Trimesh.vertex_normals`` -- both are inputs here."""
import numpy as np


# golden cases (tests/golden/dataprep.npz)
CASES = {  # name: (seed, T, V, F, S, H, W, degenerate, dtype)
    "small": (1, 3, 50, 80, 64, 16, 24, False, np.float32),
    "degenerate": (2, 2, 40, 60, 33, 8, 8, True, np.float64),
    "ragged": (3, 5, 997, 1990, 1000, 257, 129, False, np.float32),
}


def sample_texture_color_vectorized(uvs, texture_array):
    """dataset_utils.py:19-41."""
    u, v = uvs[:, 0], uvs[:, 1]
    x = (u * (texture_array.shape[1] - 1)).astype(int)          # :33
    y = ((1 - v) * (texture_array.shape[0] - 1)).astype(int)    # :34
    x = np.clip(x, 0, texture_array.shape[1] - 1)               # :36
    y = np.clip(y, 0, texture_array.shape[0] - 1)               # :37
    return texture_array[y, x], np.stack([y, x], axis=1)


def track(vertex_frames, faces, face_indices, bary, face_uvs, texture_array, vertex_normals=None):
    """dataset_utils.py:85-133 given the samples (face_indices, bary) and per-frame vertex normals."""
    bary = np.asarray(bary, dtype=np.float64)
    uvs = np.einsum("ij,ijk->ik", bary, np.asarray(face_uvs, dtype=np.float64)[face_indices])      # :88-95
    texels, yx = sample_texture_color_vectorized(uvs, texture_array)
    rgb = texels / 255.0                                                                           # :97-98
    pts, nrm = [], []
    for t in range(vertex_frames.shape[0]):
        tri = np.asarray(vertex_frames[t], dtype=np.float64)[faces[face_indices]]                  # :112
        pts.append((tri * bary.reshape((-1, 3, 1))).sum(axis=1))                                   # :113 trimesh barycentric_to_points
        if vertex_normals is not None:
            vn = np.asarray(vertex_normals[t], dtype=np.float64)[faces[face_indices]]              # :116
            n = np.einsum("ij,ijk->ik", bary, vn)                                                  # :118-122
            norms = np.linalg.norm(n, axis=1, keepdims=True)                                       # :124
            n = n / np.where(norms == 0, 1.0, norms)                                               # :125-126
            nrm.append(n)
    T = vertex_frames.shape[0]
    out_n = np.stack(nrm).astype(np.float32) if nrm else None
    return (np.stack(pts).astype(np.float32), out_n, np.tile(rgb[None], (T, 1, 1)).astype(np.float32), yx)


def make_case(seed, T, V, F, S, H, W, degenerate=False, dtype=np.float32):
    """Seeded synthetic deforming mesh + samples (the Dyscene16k data itself is unreleased, README.md:97)."""
    rng = np.random.default_rng(seed)
    verts0 = rng.uniform(-0.5, 0.5, size=(V, 3))
    frames = np.stack([verts0 + 0.05 * t * rng.normal(size=(V, 3)) for t in range(T)]).astype(dtype)
    faces = rng.integers(0, V, size=(F, 3)).astype(np.int64)
    face_indices = rng.integers(0, F, size=S).astype(np.int64)
    r = rng.uniform(size=(S, 2))
    s1 = np.sqrt(r[:, 0])
    bary = np.stack([1 - s1, s1 * (1 - r[:, 1]), s1 * r[:, 1]], axis=1)       # uniform on the triangle, float64
    vn = rng.normal(size=(T, V, 3))
    vn /= np.linalg.norm(vn, axis=2, keepdims=True)
    vn = vn.astype(dtype)
    if degenerate and S > 0:
        vn[:, faces[face_indices[0]]] = 0.0        # zero interpolated normal -> the norms == 0 branch (:125)
        bary[1 % S] = [1.0, 0.0, 0.0]              # a sample exactly on a vertex
    face_uvs = rng.uniform(-0.05, 1.05, size=(F, 3, 2))                       # slightly outside [0, 1]: exercises np.clip
    if degenerate:
        face_uvs[face_indices[2 % max(S, 1)]] = [[1.0, 0.0], [1.0, 0.0], [1.0, 0.0]]     # u = 1, v = 0 -> last texel exactly
    tex = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)
    return dict(vertex_frames=frames, faces=faces, face_indices=face_indices, bary=bary, vertex_normals=vn, face_uvs=face_uvs, texture=tex)
