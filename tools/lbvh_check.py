#!/usr/bin/env python3
"""Sponza with the device builders (lbvh: radix tree, ploc: locally-ordered clustering) vs the host SBVH: build time, tree size,
frame time (configs[1] camera, 8 spp)."""
import json, sys, time
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO / "rendering-fw_b200" / "python"))
import numpy as np
import rfwb200 as R, scenes as S
sc = S.sponza_or_standin()
out = {}
ref = None
for builder in ("sbvh", "lbvh", "lbvh-presplit", "ploc", "ploc-presplit"):
    ctx = R.RenderContext(R.load_product())
    ctx.set_setting("lbvh_presplit", "on" if builder.endswith("presplit") else "off")
    ctx.set_setting("builder", builder.split("-")[0])
    t0 = time.time(); S.upload(ctx, sc, 1920, 1080); up = time.time() - t0
    g = ctx.get_geometry_stats()
    ctx.set_setting("spp", 8)
    cam = sc.camera(1920, 1080)
    ms = []
    for _ in range(8):
        ctx.render_frame(cam, R.RESET); ctx.synchronize(); ms.append(ctx.get_stats().render_time)
    img = ctx.read_image()
    if ref is None: ref = img.copy()
    out[builder] = {"upload_s": up, "on_device": int(g.on_device), "device_ms": g.device_ms, "host_ms": g.host_ms, "bvh": ctx.get_bvh_info(),
                    "frame_ms": float(np.median(ms[2:])), "pixels_differing": float((np.abs(img - ref).max(axis=-1) > 0).mean()), "mean": float(img[..., :3].mean())}
    ctx.close()
print(json.dumps(out))
