"""ORACLE SUPPORT (test infrastructure): the recipe that puts the UNMODIFIED reference next to the oracle.

The reference is a pure-Python script tree (no package, nothing to compile), so "building" it is staging the files of the
hot path -- byte for byte, from where they lie under /root/reference -- into ``oracle/_ref/`` (git-ignored, NOT
gpurun-ignored: it travels to the GPU box like a built .so).  On the box ``bench.py --impl reference`` and the
``reference_gpu`` leg import the reference from there through the shims of oracle/ref_shims.py; nothing is ever copied
into tracked paths and every file is checked against its source by SHA-256 (``MANIFEST.json``).

    python -m oracle.build_ref            # stage (no-op when /root/reference is absent: the GPU box uses the staged files)
    python -m oracle.build_ref --verify   # re-hash the staged files against the manifest

Called by __graft_entry__.build().
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference"
DST = os.path.join(HERE, "_ref")

# the hot path, its two callers, the glue they import, the config and the config-(d) demo assets
FILES = [
    "model/__init__.py", "model/Pcd_motion.py", "model/transformer.py", "model/loss.py",
    "model/image_encoder/dinov2.py",
    "train.py", "setup.py", "configs/dyscene.yaml", "configs/api_keys.yaml",
    "utils/training_utils.py", "utils/inference_utils.py", "utils/mesh_processing.py",
    "dataset/dataset_utils.py", "dataset/dyscene.py",
    "scripts/inference_with_video_mesh.py",
    "examples/chili.glb", "examples/chili.mp4",
]
OPTIONAL = ["model/image_encoder/__init__.py", "utils/__init__.py", "dataset/__init__.py"]


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def available():
    """True when a staged (or live) reference tree can be imported."""
    return os.path.exists(os.path.join(DST, "MANIFEST.json")) or os.path.isdir(SRC)


def root():
    """Directory to put on sys.path: the live tree in the build container, the staged copy on the GPU box."""
    if os.path.isdir(os.path.join(SRC, "model")):
        return SRC
    if os.path.exists(os.path.join(DST, "MANIFEST.json")):
        return DST
    raise FileNotFoundError("no reference tree: /root/reference is absent and oracle/_ref was not staged (python -m oracle.build_ref)")


def build(verbose=False):
    if not os.path.isdir(SRC):
        return DST if os.path.exists(os.path.join(DST, "MANIFEST.json")) else None
    manifest = {}
    for rel in FILES + [f for f in OPTIONAL if os.path.exists(os.path.join(SRC, f))]:
        s, d = os.path.join(SRC, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        if not (os.path.exists(d) and os.path.getsize(d) == os.path.getsize(s) and _sha(d) == _sha(s)):
            shutil.copyfile(s, d)
            os.chmod(d, 0o644)
        manifest[rel] = _sha(d)
        if verbose:
            print(rel, manifest[rel][:12])
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC, "files": manifest}, f, indent=1, sort_keys=True)
    return DST


def verify():
    with open(os.path.join(DST, "MANIFEST.json")) as f:
        m = json.load(f)["files"]
    bad = [rel for rel, h in m.items() if _sha(os.path.join(DST, rel)) != h]
    if bad:
        raise RuntimeError(f"oracle/_ref differs from its manifest: {bad}")
    return len(m)


if __name__ == "__main__":
    if "--verify" in sys.argv:
        print("verified", verify(), "files")
    else:
        print(build(verbose=True))
