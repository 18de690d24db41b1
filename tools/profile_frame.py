#!/usr/bin/env python3
"""Render a few Sponza frames through the C ABI — the short command ncu wraps (B200_PROFILING.md)."""
import argparse
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO / "rendering-fw_b200" / "python"))
import rfwb200 as R  # noqa: E402
import scenes as S  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--spp", type=int, default=8)
ap.add_argument("--frames", type=int, default=1)
ap.add_argument("--width", type=int, default=1920)
ap.add_argument("--height", type=int, default=1080)
ap.add_argument("--smem-nodes", type=int, default=None)
ap.add_argument("--scene", default="sponza")
ap.add_argument("--set", nargs="*", default=[], help="key=value settings")
ap.add_argument("--timing", action="store_true")
a = ap.parse_args()
sc = S.sponza_or_standin() if a.scene == "sponza" else S.cornell_box(unit_scale=True)
ctx = R.RenderContext(R.load_product())
S.upload(ctx, sc, a.width, a.height)
ctx.set_setting("spp", a.spp)
if a.smem_nodes is not None:
    ctx.set_setting("smem_nodes", a.smem_nodes)
for kv in a.set:
    k, v = kv.split("=")
    ctx.set_setting(k, v)
ctx.update()
if a.timing:
    ctx.set_setting("timing", "on")
cam = sc.camera(a.width, a.height)
for _ in range(a.frames):
    ctx.render_frame(cam, R.RESET)
ctx.synchronize()
st = ctx.get_stats()
print("frame ms %.3f primary %.3f trace %.3f (d1 %.3f d2+ %.3f) shade %.3f" % (st.render_time, st.primary_time, st.secondary_time + st.deep_time, st.secondary_time, st.deep_time, st.shade_time), a.set)
