#!/usr/bin/env python3
"""BASELINE.json configs[3]: animated skinned scene, per-frame BVH refit, 1920x1080.

Per frame k (t_k = k/60 s): pose -> geometry update -> render_frame(RESET), `--spp` samples.  Two routes are timed:
  device : rfwb200_set_mesh_pose (joint matrices only) -> k_skin_vertices + k_update_triangles -> rfwb200_update =
           k_refit + k_flatten_shade; nothing but 64 B per joint crosses PCIe
  host   : the reference's route through the plugin boundary: the caller skins on the CPU (here: numpy restatement of
           gltf/mesh.cpp set_pose, timed separately), set_mesh re-sends the mesh, update() flattens + refits on the host
           and re-uploads every record (setting refit=host)
Prints one JSON line with the frame-time split (skin / refit / trace) of both routes.
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO / "rendering-fw_b200" / "python"))
sys.path.insert(0, str(REPO))
import rfwb200 as R  # noqa: E402
import scenes as S  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=120)
ap.add_argument("--spp", type=int, default=1)
ap.add_argument("--width", type=int, default=1920)
ap.add_argument("--height", type=int, default=1080)
ap.add_argument("--copies", type=int, default=1)
ap.add_argument("--with-sponza", action="store_true", help="put the animated mesh inside the config-2 scene (262k static triangles)")
a = ap.parse_args()


def build_scene():
    sc, skins = S.animated_config4(a.copies)
    if a.with_sponza:
        sp = S.sponza_or_standin()
        # append the skinned meshes to Sponza (materials / textures re-indexed)
        mat_off, tex_off, mesh_off = len(sp.materials), len(sp.textures), len(sp.meshes)
        mats = sc.materials.copy()  # texaddr is patched by set_materials from tex_ids
        ids = sc.tex_ids.copy()
        ids[ids >= 0] += tex_off
        sp.materials = np.concatenate([sp.materials, mats])
        sp.tex_ids = np.concatenate([sp.tex_ids, ids])
        sp.textures = list(sp.textures) + list(sc.textures)
        new_skins = []
        for sk in skins:
            m = sc.meshes[sk.mesh_index]
            m.triangles["material"] += mat_off
            inst = [M for mi, M in sc.instances if mi == sk.mesh_index][0]
            sp.meshes.append(m)
            sp.instances.append((len(sp.meshes) - 1, S.translate(*sp.camera_pos) @ S.translate(6.0, -9.5, 0.3) @ S.scale(4.0) @ inst))
            sk.mesh_index = len(sp.meshes) - 1
            new_skins.append(sk)
        return sp, new_skins
    return sc, skins


def run(route: str):
    from oracle import skinning as K  # CPU restatement of the reference's skinning: the host route's caller-side work

    sc, skins = build_scene()
    ctx = R.RenderContext(R.load_product())
    ctx.set_setting("refit", route)
    S.upload(ctx, sc, a.width, a.height)
    ctx.set_setting("spp", a.spp)
    if route == "device":
        for sk in skins:
            ctx.set_mesh_skin(sk.mesh_index, sk.base_vertices, sk.base_normals, sk.joints, sk.weights)
    cam = sc.camera(a.width, a.height)
    t_skin = t_update = t_frame = t_render_dev = t_geo_dev = 0.0
    for k in range(-3, a.frames):  # 3 warm-up frames
        t0 = time.perf_counter()
        if route == "device":
            for sk in skins:
                ctx.set_mesh_pose(sk.mesh_index, sk.joint_matrices(k))
            t1 = time.perf_counter()
        else:
            posed = []
            for sk in skins:
                m = sc.meshes[sk.mesh_index]
                v, n = K.set_pose(sk.base_vertices, sk.base_normals, sk.joints, sk.weights, sk.joint_matrices(k))
                posed.append((sk.mesh_index, v, K.update_triangles(m.triangles, v, n, m.indices), m.indices))
            t1 = time.perf_counter()
            for mi, v, t, idx in posed:
                ctx.set_mesh(mi, v, t, idx)
        ctx.update()
        t2 = time.perf_counter()
        ctx.render_frame(cam, R.RESET)
        ctx.synchronize()
        t3 = time.perf_counter()
        if k >= 0:
            t_skin += t1 - t0
            t_update += t2 - t1
            t_frame += t3 - t0
            t_render_dev += ctx.get_stats().render_time
            g = ctx.get_geometry_stats()
            t_geo_dev += g.device_ms if g.on_device else 0.0
    n = a.frames
    g = ctx.get_geometry_stats()
    info = ctx.get_bvh_info()
    img = ctx.read_image()
    out = {"route": route, "frame_ms": 1e3 * t_frame / n, "pose_or_cpu_skin_ms": 1e3 * t_skin / n, "update_call_ms": 1e3 * t_update / n,
           "render_device_ms": t_render_dev / n, "geometry_device_ms": t_geo_dev / n, "refits": int(g.refits), "builds": int(g.builds),
           "fps": n / t_frame, "image_mean": float(img[..., :3].mean()), "triangles": info["triangles"], "nodes": info["nodes"]}
    ctx.close()
    return out


res = {"config": f"{'sponza + ' if a.with_sponza else ''}{S.animated_config4(1)[0].name} x{a.copies}, {a.width}x{a.height}, {a.spp} spp, {a.frames} frames",
       "device": run("device"), "host": run("host")}
res["msamples_per_s_incl_refit"] = {r: a.width * a.height * a.spp / (res[r]["frame_ms"] * 1e3) for r in ("device", "host")}
print(json.dumps(res))
