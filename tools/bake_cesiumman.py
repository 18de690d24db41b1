#!/usr/bin/env python3
"""Bake BASELINE.json configs[3]'s asset — assets/models/CesiumMan (4,672 triangles, 3,273 vertices, 19 joints, one
57-channel animation) — into rendering-fw_b200/data/_baked/cesiumman.npz: bind pose, joints/weights, texture and the
joint matrices of frames t_k = k/60 s, k = 0..119.  The reference does this work in its glTF
loader and animation system, which are callers of the plugin boundary (out of scope as product code); this tool
restates just enough of them to feed the path:
  * accessors / skin / node hierarchy      RFW/system/src/rfw/geometry/gltf/object.cpp, hierarcy.cpp
  * sampling: LINEAR = lerp (nlerp for rotations), f <= 0 -> first key       gltf/animation.cpp:236-310
  * node transform T * R * S * matrix, combined = parent * local              gltf/node.cpp:55-64,113-121
  * joint matrix = inverse(meshNode.combined) * jointNode.combined * inverseBind[j]   gltf/node.cpp:97-104
The baked file is git-ignored (reference data is never committed) but travels to the GPU box.
"""
import json
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
SRC = Path("/root/reference/assets/models/CesiumMan")
OUT = REPO / "rendering-fw_b200" / "data" / "_baked" / "cesiumman.npz"
FRAMES, FPS = 120, 60.0

CT = {5120: np.int8, 5121: np.uint8, 5122: np.int16, 5123: np.uint16, 5125: np.uint32, 5126: np.float32}
NC = {"SCALAR": 1, "VEC2": 2, "VEC3": 3, "VEC4": 4, "MAT4": 16}


def accessor(g, buf, i):
    a = g["accessors"][i]
    bv = g["bufferViews"][a["bufferView"]]
    dt, n = np.dtype(CT[a["componentType"]]), NC[a["type"]]
    off = bv.get("byteOffset", 0) + a.get("byteOffset", 0)
    stride = bv.get("byteStride", 0) or dt.itemsize * n
    raw = np.frombuffer(buf, np.uint8, count=stride * (a["count"] - 1) + dt.itemsize * n, offset=off)
    out = np.zeros((a["count"], n), dt)
    for k in range(a["count"]):
        out[k] = np.frombuffer(raw[k * stride:k * stride + dt.itemsize * n].tobytes(), dt)
    return out


def quat_to_mat(q):  # glm::mat4(quat), q = (x, y, z, w) as stored in glTF
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w), 0],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w), 0],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y), 0],
                     [0, 0, 0, 1]], np.float64)


def main():
    if not (SRC / "CesiumMan.gltf").exists():
        print("CesiumMan asset not found; nothing baked")
        return 1
    g = json.loads((SRC / "CesiumMan.gltf").read_text())
    buf = (SRC / g["buffers"][0]["uri"]).read_bytes()
    prim = g["meshes"][0]["primitives"][0]
    at = prim["attributes"]
    pos, nrm, uv = accessor(g, buf, at["POSITION"]), accessor(g, buf, at["NORMAL"]), accessor(g, buf, at["TEXCOORD_0"])
    joints, weights = accessor(g, buf, at["JOINTS_0"]).astype(np.uint32), accessor(g, buf, at["WEIGHTS_0"]).astype(np.float32)
    weights = weights / weights.sum(1, keepdims=True)  # the loader normalises them (gltf/object.cpp:434)
    idx = accessor(g, buf, prim["indices"]).astype(np.uint32).reshape(-1, 3)
    skin = g["skins"][0]
    ibm = accessor(g, buf, skin["inverseBindMatrices"]).reshape(-1, 4, 4).transpose(0, 2, 1).astype(np.float64)  # column-major -> (row, col)
    nodes = g["nodes"]
    parent = {c: i for i, n in enumerate(nodes) for c in n.get("children", [])}
    mesh_node = [i for i, n in enumerate(nodes) if "mesh" in n][0]
    anim = g["animations"][0]
    samplers = [(accessor(g, buf, s["input"])[:, 0].astype(np.float64), accessor(g, buf, s["output"]).astype(np.float64), s.get("interpolation", "LINEAR"))
                for s in anim["samplers"]]

    def sample(si, t, is_quat):
        keys, vals, method = samplers[si]
        dur = keys[-1]
        if t > dur:
            t = np.fmod(t, dur)
        k = int(np.clip(np.searchsorted(keys, t, side="right") - 1, 0, len(keys) - 2))
        f = (t - keys[k]) / (keys[k + 1] - keys[k])
        if f <= 0:
            v = vals[0]
        elif method == "STEP":
            v = vals[k]
        else:
            v = (1 - f) * vals[k] + f * vals[k + 1]
        return v / np.linalg.norm(v) if is_quat else v

    def local(i, trs):
        n = nodes[i]
        T, Rq, Sc = trs.get((i, "translation"), n.get("translation", [0, 0, 0])), trs.get((i, "rotation"), n.get("rotation", [0, 0, 0, 1])), \
            trs.get((i, "scale"), n.get("scale", [1, 1, 1]))
        M = np.array(n["matrix"], np.float64).reshape(4, 4).T if "matrix" in n else np.eye(4)
        Tm = np.eye(4)
        Tm[:3, 3] = T
        return Tm @ quat_to_mat(Rq) @ np.diag(list(Sc) + [1.0]) @ M

    def combined(i, trs, cache):
        if i not in cache:
            L = local(i, trs)
            cache[i] = combined(parent[i], trs, cache) @ L if i in parent else L
        return cache[i]

    poses = np.zeros((FRAMES, len(skin["joints"]), 4, 4), np.float32)
    mesh_xf = None
    for k in range(FRAMES):
        t = k / FPS
        trs = {(c["target"]["node"], c["target"]["path"]): sample(c["sampler"], t, c["target"]["path"] == "rotation") for c in anim["channels"]}
        cache = {}
        Mm = combined(mesh_node, trs, cache)
        if mesh_xf is None:
            mesh_xf = Mm
        inv = np.linalg.inv(Mm)
        for j, jn in enumerate(skin["joints"]):
            poses[k, j] = (inv @ combined(jn, trs, cache) @ ibm[j]).astype(np.float32)
    tex = np.zeros((0, 0, 4), np.uint8)
    try:
        from PIL import Image

        im = Image.open(SRC / g["images"][0]["uri"]).convert("RGBA").resize((512, 512))
        tex = np.asarray(im, np.uint8)[::-1].copy()  # rows bottom-up like the reference's FreeImage load
    except Exception as e:  # noqa: BLE001
        print("texture skipped:", e)
    OUT.parent.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(OUT, positions=pos.astype(np.float32), normals=nrm.astype(np.float32), uvs=uv.astype(np.float32), indices=idx,
                        joints=joints, weights=weights, poses=poses, mesh_transform=mesh_xf.astype(np.float64), texture=tex)
    print(f"baked {OUT}: {len(idx)} triangles, {len(pos)} vertices, {poses.shape[1]} joints, {FRAMES} frames, texture {tex.shape}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
