#!/usr/bin/env python3
"""Bake the reference's Sponza setup (BASELINE.json configs[1]) into the wire
formats of the plugin boundary, so the GPU box — where /root/reference does not exist — can render it.

Reads   /root/reference/assets/models/sponza/{sponza.obj,sponza.mtl,textures/*.tga}
        /root/reference/assets/envmaps/sky_15.hdr
Writes  rendering-fw_b200/data/_baked/sponza.rfwscene   (git-ignored, travels with gpurun snapshots)

It restates what the reference's loaders do on the way to the boundary (all out of scope as product code):
  one rfw::Mesh per OBJ object/material with mesh-local indices   geometry/assimp/object.cpp:351-751
  MTL -> HostMaterial                                             material_list.cpp:34-177
  TGA -> RGBA8 + 5 box-filtered mips, HAS_ALPHA when alpha hits 0  texture.cpp:16-126,163-225
  per-triangle LOD constant                                       assimp/object.cpp:727-731
  scene: sponza at scale 0.2, 20x100 quad light (100,100,100) at y = 60, sky_15.hdr
                                                                  Examples/imgui_app/main.cpp:88-110,147-149
No reference source code is copied; only asset data is converted.
"""
import argparse
import sys
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO / "rendering-fw_b200" / "python"))
import rfwb200 as R  # noqa: E402
import scenes as S  # noqa: E402

ASSETS = Path("/root/reference/assets")


def parse_mtl(path: Path):
    mats, cur = {}, None
    for line in path.read_text().splitlines():
        p = line.split()
        if not p or p[0].startswith("#"):
            continue
        if p[0] == "newmtl":
            cur = {"name": p[1], "Kd": (0.0, 0.0, 0.0), "Ke": (0.0, 0.0, 0.0), "Ns": 0.0, "Ni": 1.0, "d": None}
            mats[p[1]] = cur
        elif cur is not None:
            if p[0] in ("Kd", "Ke"):
                cur[p[0]] = tuple(float(x) for x in p[1:4])
            elif p[0] in ("Ns", "Ni", "d"):
                cur[p[0]] = float(p[1])
            elif p[0] in ("map_Kd", "map_d", "map_Disp", "map_bump", "bump"):
                cur[p[0]] = p[-1]
    return mats


def load_tga(path: Path, max_dim: int):
    from PIL import Image

    im = Image.open(path)
    has_alpha = im.mode in ("RGBA", "LA")
    im = im.convert("RGBA")
    while max(im.size) > max_dim:
        im = im.resize((im.size[0] // 2, im.size[1] // 2), Image.BOX)
    a = np.asarray(im, dtype=np.uint8)[::-1].copy()  # row 0 = bottom row, like FreeImage
    alpha_flag = bool(has_alpha and (a[..., 3] == 0).any())  # histogram[0] > 0, texture.cpp:57-60
    return a, alpha_flag


def load_hdr(path: Path, out_w: int, out_h: int):
    """(0, 0): the asset as shipped — its RGBE texels are kept and decoded like FreeImage does (scenes.decode_rgbe); any other
    size: area-resampled with OpenCV (smaller snapshot, stated in the scene name)."""
    if out_w == 0:
        rgbe = S.read_radiance_hdr(path)
        h, w = rgbe.shape[:2]
        return S.decode_rgbe(rgbe).reshape(-1, 3), w, h, "sky_15.hdr %dx%d as shipped" % (w, h), rgbe
    try:
        import cv2

        img = cv2.imread(str(path), cv2.IMREAD_UNCHANGED)
        if img is None:
            raise RuntimeError("cv2 could not read the HDR")
        img = img[..., ::-1].astype(np.float32)  # BGR -> RGB
        img = cv2.resize(img, (out_w, out_h), interpolation=cv2.INTER_AREA)
        return img.reshape(-1, 3), out_w, out_h, "sky_15.hdr (area-resampled to %dx%d)" % (out_w, out_h), None
    except Exception as e:  # constant sky, as SURVEY.md §8d allows — stated in the scene name
        print("HDR decode unavailable (%s): constant sky (0.5,0.6,0.8)" % e, file=sys.stderr)
        return np.tile(np.array([[0.5, 0.6, 0.8]], np.float32), (1, 1)), 1, 1, "constant sky", None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-dim", type=int, default=4096, help="textures larger than this are box-downsampled (the asset's largest are 2048: default = as shipped)")
    ap.add_argument("--sky", type=int, nargs=2, default=(0, 0), help="resample the sky to this size; 0 0 = as shipped")
    ap.add_argument("--out", type=Path, default=S.BAKED_SPONZA)
    args = ap.parse_args()
    obj = ASSETS / "models" / "sponza" / "sponza.obj"
    if not obj.exists():
        print("reference assets not present; nothing baked", file=sys.stderr)
        return 0
    t0 = time.time()
    mtl = parse_mtl(obj.with_suffix(".mtl"))
    scene = S.Scene(name="sponza")

    # ---- textures + materials -------------------------------------------------------------------------
    tex_index, tex_alpha = {}, {}
    mat_index = {}
    for name, m in mtl.items():
        tex0 = -1
        has_alpha = False
        if "map_Kd" in m:
            f = m["map_Kd"]
            if f not in tex_index:
                rgba, aflag = load_tga(obj.parent / f, args.max_dim)
                tex_index[f] = S.add_texture_rgba8(scene, rgba)
                tex_alpha[f] = aflag
            tex0, has_alpha = tex_index[f], tex_alpha[f]
        # what assimp's OBJ importer hands to the reference's material rule: Ke, Kd, Tf (absent: 0), d (absent: 0 = "not set"),
        # Ns, no shininess strength, Ni, no reflectivity
        r = S.material_rule(m["Ke"], m["Kd"], (0.0, 0.0, 0.0), m["d"] if m["d"] is not None else 0.0, m["Ns"], 0.0, m["Ni"], 0.0)
        color = tuple(float(c) for c in r["color"])
        if tex0 >= 0 and all(c == 0 for c in color):
            color = (1.0, 1.0, 1.0)  # material_list.cpp:84-86: a black diffuse colour under a diffuse map becomes white
        rough, transmission, eta = float(r["roughness"]), float(r["transmission"]), float(r["eta"])
        mat_index[name] = S.add_material(scene, color, roughness=rough, transmission=transmission, eta=eta, tex0=tex0,
                                         smooth=True, has_alpha=has_alpha)
    light_mat = S.add_material(scene, (100, 100, 100), roughness=1.0)
    print("textures %d materials %d (%.1fs)" % (len(scene.textures), len(scene.materials), time.time() - t0))

    # ---- geometry ----------------------------------------------------------------------------------------
    V, VT, VN = [], [], []
    groups = []  # (material name, list of faces (9 ints))
    cur_faces, cur_mat = None, None
    with open(obj) as fh:
        for line in fh:
            if line.startswith("v "):
                V.append(line.split()[1:4])
            elif line.startswith("vt "):
                VT.append(line.split()[1:3])
            elif line.startswith("vn "):
                VN.append(line.split()[1:4])
            elif line.startswith("usemtl"):
                cur_mat = line.split()[1]
                cur_faces = []
                groups.append((cur_mat, cur_faces))
            elif line.startswith("f "):
                p = line.split()[1:]
                idx = [[int(x) if x else 0 for x in q.split("/")] for q in p]
                for k in range(1, len(idx) - 1):  # aiProcess_Triangulate (fan)
                    cur_faces.append(idx[0] + idx[k] + idx[k + 1])
    V, VT, VN = np.array(V, np.float32), np.array(VT, np.float32), np.array(VN, np.float32)
    total = 0
    for mat_name, faces in groups:
        if not faces:
            continue
        f = np.array(faces, np.int64).reshape(-1, 3, 3) - 1  # (nt, corner, v/vt/vn)
        keys = f.reshape(-1, 3)
        uniq, inv = np.unique(keys, axis=0, return_inverse=True)  # aiProcess_JoinIdenticalVertices
        verts = V[uniq[:, 0]]
        uvs = VT[uniq[:, 1]] if len(VT) else np.zeros((len(uniq), 2), np.float32)
        nrm = VN[uniq[:, 2]]
        nrm = nrm / np.maximum(np.linalg.norm(nrm, axis=1, keepdims=True), 1e-20)
        idx = inv.reshape(-1, 3).astype(np.uint32)
        mi = mat_index[mat_name]
        tdims = None
        t0id = int(scene.tex_ids[mi, 0])
        if t0id >= 0:
            tdims = (scene.textures[t0id]["width"], scene.textures[t0id]["height"])
        tri = S.make_triangles(verts[idx], nrm[idx], uvs[idx], mi, tdims)
        scene.meshes.append(S.SceneMesh(np.concatenate([verts, np.ones((len(verts), 1), np.float32)], 1), tri, idx))
        total += len(idx)
    print("meshes %d triangles %d (%.1fs)" % (len(scene.meshes), total, time.time() - t0))
    k = S.scale(0.2)
    scene.instances = [(i, k) for i in range(len(scene.meshes))]
    scene.meshes.append(S.quad((0, -1, 0), (0, 0, 0), 20.0, 100.0, light_mat))
    scene.instances.append((len(scene.meshes) - 1, S.translate(0, 60.0, 0)))

    sky, sw, sh, sky_note, rgbe = load_hdr(ASSETS / "envmaps" / "sky_15.hdr", *args.sky)
    scene.sky = (sky, sw, sh)
    scene.sky_rgbe = rgbe
    scene.camera_pos, scene.camera_dir, scene.fov = (0.0, 10.0, 0.0), (1.0, 0.0, 0.05), 40.0
    tex_note = "textures as shipped" if args.max_dim >= 2048 else "textures<=%d" % args.max_dim
    scene.name = "sponza (262k tris, %d meshes, %s, %s)" % (len(scene.meshes), tex_note, sky_note)
    S.save_baked(scene, args.out)
    print("wrote %s: %.1f MiB in %.1fs" % (args.out, args.out.stat().st_size / 2 ** 20, time.time() - t0))
    return 0


if __name__ == "__main__":
    sys.exit(main())
