"""Golden fixture for SURVEY.md 8(f3), produced by the reference's OWN functions: compute_chamfer_distance and
compute_fscore are extracted from /root/reference/evaluation/evaluation_pcd.py with `ast` (the module imports trimesh /
matplotlib, absent here) and executed unmodified on seeded point sets.

    python tests/golden/make_golden_chamfer.py
"""
import ast
import os

import numpy as np
from scipy.spatial import cKDTree

REF = "/root/reference/evaluation/evaluation_pcd.py"


def extract(name):
    tree = ast.parse(open(REF).read())
    node = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == name)
    ns = dict(np=np, cKDTree=cKDTree)
    exec(compile(ast.Module([node], []), REF, "exec"), ns)
    return ns[name]


def main():
    cd, fs = extract("compute_chamfer_distance"), extract("compute_fscore")
    rng = np.random.default_rng(11)
    out = {}
    # (n1, n2, noise): a surface-like cloud and a perturbed re-sampling of it, sizes ragged on purpose
    for k, (n1, n2, noise) in enumerate([(700, 700, 0.01), (1000, 333, 0.02), (257, 1025, 0.05), (1, 40, 0.3), (2000, 2000, 0.0)]):
        base = rng.normal(size=(max(n1, n2), 3))
        base /= np.linalg.norm(base, axis=1, keepdims=True)
        p1 = 0.5 * base[:n1] + rng.normal(size=(n1, 3)) * 0.002
        p2 = 0.5 * base[rng.permutation(max(n1, n2))[:n2]] + rng.normal(size=(n2, 3)) * noise
        out[f"c{k}_p1"], out[f"c{k}_p2"] = p1, p2
        out[f"c{k}_chamfer"] = np.float64(cd(p1, p2))
        out[f"c{k}_fscore"] = np.array([fs(p1, p2, threshold=t) for t in (0.02, 0.05, 0.001)], dtype=np.float64)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "chamfer.npz"), **out)
    print({k: v for k, v in out.items() if not k.endswith(("p1", "p2"))})


if __name__ == "__main__":
    main()
