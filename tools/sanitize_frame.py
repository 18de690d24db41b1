#!/usr/bin/env python3
"""The short command compute-sanitizer wraps: a Cornell box 128x128, 4 spp, PT-mode frame + one Converge frame +
an E-mode frame through the C ABI, single context and a 2-rank in-process group (peer stores, flow-control kernels)."""
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
print("sanitize_frame ok")
