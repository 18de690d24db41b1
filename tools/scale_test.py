#!/usr/bin/env python3
"""BASELINE.json configs[2] stand-in (San Miguel is not available offline): a synthetic multi-million-triangle atrium,
1920x1080 4 spp — exercises the flattened-BVH builder and deep traversal at scale."""
import sys, time
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO / "rendering-fw_b200" / "python"))
import numpy as np
import rfwb200 as R, scenes as S
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
builder = sys.argv[2] if len(sys.argv) > 2 else "sbvh"  # sbvh = host SAH builder with spatial splits, lbvh = device builder
t0 = time.time(); sc = S.atrium(target_tris=n); print("scene", sc.name, sc.triangle_count(), "tris, gen %.1fs" % (time.time() - t0))
ctx = R.RenderContext(R.load_product())
ctx.set_setting("builder", builder)
t0 = time.time(); S.upload(ctx, sc, 1920, 1080); print(builder, "upload+build %.1fs" % (time.time() - t0), ctx.get_bvh_info())
g = ctx.get_geometry_stats(); print("geometry: on_device", g.on_device, "device_ms %.2f host_ms %.1f" % (g.device_ms, g.host_ms))
ctx.set_setting("spp", 4)
cam = sc.camera(1920, 1080)
for i in range(3):
    ctx.render_frame(cam, R.RESET); ctx.synchronize()
    st = ctx.get_stats()
    print("frame ms %.2f -> %.1f Msamples/s" % (st.render_time, 1920 * 1080 * 4 / st.render_time / 1e3))
img = ctx.read_image(); print("mean", img[..., :3].mean(), "finite", np.isfinite(img).all())
