"""Golden fixture for SURVEY.md 8(f4), produced by the reference's OWN code: /root/reference/dataset/dataset_utils.py is
imported unmodified with a stand-in for the absent ``trimesh`` package.  The stand-in only REPLAYS fixed samples
(sample_surface -> the seeded (points, face_indices) of the case; points_to_barycentric -> the seeded barycentric
coordinates; Trimesh(...).vertex_normals -> the seeded per-frame normals; barycentric_to_points = trimesh's published
one-liner), so every line of track_with_normal_rgb / sample_texture_color_vectorized itself runs as written.

    python tests/golden/make_golden_dataprep.py
"""
import importlib.util
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import dataprep_oracle as orc  # noqa: E402

CASES = orc.CASES


def load_reference(case):
    tm = types.ModuleType("trimesh")

    class Trimesh:
        def __init__(self, vertices=None, faces=None, process=False):
            self._v, self.faces = vertices, faces

        @property
        def vertices(self):
            return self._v

        @vertices.setter
        def vertices(self, v):
            self._v = np.asarray(v, dtype=np.float64)        # trimesh stores vertices as float64
            self._t = next(t for t in range(case["vertex_frames"].shape[0]) if np.array_equal(case["vertex_frames"][t], v))

        @property
        def vertex_normals(self):
            return np.asarray(case["vertex_normals"][self._t], dtype=np.float64)

    tm.Trimesh = Trimesh
    tm.sample = types.SimpleNamespace(sample_surface=lambda mesh, n: (None, case["face_indices"]))
    tm.triangles = types.SimpleNamespace(
        points_to_barycentric=lambda triangles, points: case["bary"],
        barycentric_to_points=lambda triangles, barycentric: (triangles * np.asarray(barycentric).reshape((-1, 3, 1))).sum(axis=1))
    sys.modules["trimesh"] = tm
    spec = importlib.util.spec_from_file_location("ref_dataset_utils", "/root/reference/dataset/dataset_utils.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod, Trimesh


def main():
    out = {}
    for name, args in CASES.items():
        case = orc.make_case(*args)
        mod, Trimesh = load_reference(case)
        init = Trimesh(vertices=np.asarray(case["vertex_frames"][0], dtype=np.float64), faces=case["faces"])
        pts, nrm, rgb, fidx = mod.track_with_normal_rgb(init, case["vertex_frames"], case["faces"], len(case["face_indices"]),
                                                       case["face_uvs"], case["texture"])
        out[name + "_points"], out[name + "_normals"], out[name + "_rgbs"] = pts.numpy(), nrm.numpy(), rgb.numpy()
        uvs = np.random.default_rng(args[0] + 100).uniform(-0.2, 1.2, size=(500, 2))
        out[name + "_texels"] = mod.sample_texture_color_vectorized(uvs, case["texture"])
        print(name, pts.shape, float(np.abs(pts.numpy()).mean()), float(rgb.numpy().mean()))
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "dataprep.npz"), **out)


if __name__ == "__main__":
    main()
