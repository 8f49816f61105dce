"""Minimal binary-glTF (.glb) reader for the inference entry point (trimesh is absent in this image).

Reads what /root/reference/scripts/inference_with_video_mesh.py:60-129 takes from ``trimesh.load(path, force='mesh')``: the
triangles of every mesh primitive (POSITION, NORMAL, TEXCOORD_0, indices; node transforms applied; primitives
concatenated) and the base-colour texture of the first material that has one (JPEG / PNG decoded with OpenCV).
Host code: file parsing, no arithmetic of the hot path.  trimesh post-processing that is NOT reproduced (parity unpinned,
DESIGN.md): vertex merging (``process=True``) and ``fix_normals`` winding repair -- the accessors are used as stored.
"""
import json
import struct

import numpy as np

_COMPONENT = {5120: np.int8, 5121: np.uint8, 5122: np.int16, 5123: np.uint16, 5125: np.uint32, 5126: np.float32}
_NCOMP = {"SCALAR": 1, "VEC2": 2, "VEC3": 3, "VEC4": 4, "MAT4": 16}


class GlbError(ValueError):
    pass


def _chunks(blob):
    if len(blob) < 20:
        raise GlbError("not a GLB file: shorter than its header")
    magic, version, length = struct.unpack_from("<4sII", blob, 0)
    if magic != b"glTF" or version != 2:
        raise GlbError(f"not a glTF 2.0 binary (magic {magic!r}, version {version})")
    if length > len(blob):
        raise GlbError(f"truncated GLB: header says {length} bytes, file has {len(blob)}")
    off, doc, binary = 12, None, None
    while off + 8 <= length:
        clen, ctype = struct.unpack_from("<II", blob, off)
        off += 8
        if off + clen > length:
            raise GlbError("truncated GLB chunk")
        if ctype == 0x4E4F534A:      # 'JSON'
            doc = json.loads(blob[off:off + clen].decode("utf-8"))
        elif ctype == 0x004E4942:    # 'BIN\0'
            binary = blob[off:off + clen]
        off += clen + (-clen % 4)
    if doc is None:
        raise GlbError("GLB without a JSON chunk")
    return doc, binary


def _accessor(doc, binary, index):
    acc = doc["accessors"][index]
    if "sparse" in acc:
        raise GlbError("sparse accessors are not supported")
    view = doc["bufferViews"][acc["bufferView"]]
    if view.get("buffer", 0) != 0 or binary is None:
        raise GlbError("only the embedded binary buffer is supported")
    dt = np.dtype(_COMPONENT[acc["componentType"]])
    ncomp = _NCOMP[acc["type"]]
    start = view.get("byteOffset", 0) + acc.get("byteOffset", 0)
    stride = view.get("byteStride", 0) or dt.itemsize * ncomp
    count = acc["count"]
    if start + (count - 1) * stride + dt.itemsize * ncomp > len(binary):
        raise GlbError("accessor reaches past the end of the buffer")
    if stride == dt.itemsize * ncomp:
        arr = np.frombuffer(binary, dtype=dt, count=count * ncomp, offset=start).reshape(count, ncomp)
    else:
        raw = np.frombuffer(binary, dtype=np.uint8, count=(count - 1) * stride + dt.itemsize * ncomp, offset=start)
        idx = (np.arange(count)[:, None] * stride + np.arange(dt.itemsize * ncomp)[None]).reshape(-1)
        arr = raw[idx].view(dt).reshape(count, ncomp)
    if acc.get("normalized") and dt.kind in "iu":
        arr = arr.astype(np.float32) / np.iinfo(dt).max
    return arr[:, 0] if ncomp == 1 else arr


def _node_matrix(node):
    if "matrix" in node:
        return np.array(node["matrix"], dtype=np.float64).reshape(4, 4).T
    m = np.eye(4)
    if "scale" in node:
        m = np.diag(list(node["scale"]) + [1.0]) @ m
    if "rotation" in node:
        x, y, z, w = node["rotation"]
        r = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w), 0],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w), 0],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y), 0], [0, 0, 0, 1]])
        m = r @ m
    if "translation" in node:
        t = np.eye(4)
        t[:3, 3] = node["translation"]
        m = t @ m
    return m


def _texture(doc, binary, material):
    try:
        tex = doc["textures"][material["pbrMetallicRoughness"]["baseColorTexture"]["index"]]
        image = doc["images"][tex["source"]]
        view = doc["bufferViews"][image["bufferView"]]
    except (KeyError, IndexError, TypeError):
        return None
    import cv2
    start = view.get("byteOffset", 0)
    data = np.frombuffer(binary, dtype=np.uint8, count=view["byteLength"], offset=start)
    bgr = cv2.imdecode(data, cv2.IMREAD_COLOR)
    if bgr is None:
        raise GlbError("base-colour texture could not be decoded")
    return np.ascontiguousarray(bgr[:, :, ::-1])      # RGB uint8 [H, W, 3]


def load_glb(path):
    """-> dict(vertices [V,3] f64, faces [F,3] i64, normals [V,3] f64 | None, uv [V,2] f64 | None, texture [H,W,3] u8 | None)."""
    with open(path, "rb") as f:
        doc, binary = _chunks(f.read())
    scene = doc["scenes"][doc.get("scene", 0)] if doc.get("scenes") else {"nodes": list(range(len(doc.get("nodes", []))))}
    verts, norms, uvs, faces, texture = [], [], [], [], None
    base = 0

    def visit(ni, parent):
        nonlocal base, texture
        node = doc["nodes"][ni]
        world = parent @ _node_matrix(node)
        if "mesh" in node:
            for prim in doc["meshes"][node["mesh"]]["primitives"]:
                if prim.get("mode", 4) != 4:
                    continue                     # triangles only
                at = prim["attributes"]
                p = _accessor(doc, binary, at["POSITION"]).astype(np.float64)
                p = p @ world[:3, :3].T + world[:3, 3]
                n = _accessor(doc, binary, at["NORMAL"]).astype(np.float64) if "NORMAL" in at else None
                if n is not None:
                    n = n @ np.linalg.inv(world[:3, :3])
                uv = _accessor(doc, binary, at["TEXCOORD_0"]).astype(np.float64) if "TEXCOORD_0" in at else None
                idx = _accessor(doc, binary, prim["indices"]).astype(np.int64) if "indices" in prim else np.arange(len(p), dtype=np.int64)
                if idx.size % 3 or (idx.size and (idx.min() < 0 or idx.max() >= len(p))):
                    raise GlbError("index accessor does not describe triangles of this primitive")
                verts.append(p)
                norms.append(n)
                uvs.append(uv)
                faces.append(idx.reshape(-1, 3) + base)
                base += len(p)
                if texture is None and "material" in prim:
                    texture = _texture(doc, binary, doc["materials"][prim["material"]])
        for c in node.get("children", []):
            visit(c, world)

    for ni in scene.get("nodes", []):
        visit(ni, np.eye(4))
    if not verts:
        raise GlbError("no triangle primitives in the default scene")
    V = np.concatenate(verts)
    N = np.concatenate(norms) if all(n is not None for n in norms) else None
    UV = np.concatenate(uvs) if all(u is not None for u in uvs) else None
    return dict(vertices=V, faces=np.concatenate(faces), normals=N, uv=UV, texture=texture)
