#!/usr/bin/env python3
"""The short command compute-sanitizer wraps: a Cornell box 128x128, 4 spp, PT-mode frame + one Converge frame +
an E-mode frame through the C ABI, single context and a 2-rank in-process group (peer stores, flow-control kernels); then the
feature soup (textures, all light types, instancing) as a two-level scene (k_wavefront_trace_tl, instance normals at shading
time) and on trees built by the device builders (lbvh, ploc: clustering rounds, depth-first re-ordering, collapse) with a
device refit behind them."""
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO / "rendering-fw_b200" / "python"))
import rfwb200 as R  # noqa: E402
import scenes as S  # noqa: E402

lib = R.load_product()
for devs in (None, [0, 0]):
    sc = S.cornell_box(unit_scale=True)
    ctx = R.RenderContext(lib) if devs is None else R.RenderContext(lib, devices=devs)
    S.upload(ctx, sc, 128, 128)
    ctx.set_setting("spp", 4)
    cam = sc.camera(128, 128)
    ctx.render_frame(cam, R.RESET)
    ctx.render_frame(cam, R.CONVERGE)
    img = ctx.read_image()
    print("pt", devs, float(img[..., :3].mean()))
    if devs is None:
        ctx.set_setting("mode", "embree")
        ctx.render_frame(cam, R.RESET)
        print("emode", float(ctx.read_image()[..., :3].mean()))
    ctx.close()
for label, settings in (("two-level", {"levels": 2}), ("lbvh", {"builder": "lbvh"}), ("ploc", {"builder": "ploc"}), ("ploc+presplit", {"builder": "ploc", "lbvh_presplit": "on"})):
    sc = S.feature_soup(600)
    ctx = R.RenderContext(lib)
    for k, v in settings.items():
        ctx.set_setting(k, v)
    S.upload(ctx, sc, 96, 64)
    ctx.set_setting("spp", 2)
    cam = sc.camera(96, 64)
    ctx.render_frame(cam, R.RESET)
    img = ctx.read_image()
    print(label, float(img[..., :3].mean()))
    mesh, M = sc.instances[1]
    ctx.set_instance(1, mesh, S.translate(0.05, 0.0, 0.02) @ M)  # two-level: top-level rebuild; device-built trees: device refit
    ctx.update()
    ctx.render_frame(cam, R.RESET)
    if label == "two-level":
        ctx.set_setting("mode", "embree")
        ctx.render_frame(cam, R.RESET)
    print(label, "after a moved instance", float(ctx.read_image()[..., :3].mean()))
    ctx.close()
print("sanitize_frame ok")
