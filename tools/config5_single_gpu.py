#!/usr/bin/env python3
"""BASELINE.json configs[4] on ONE GPU (the 8-GPU run shards the same frame by tiles, bench.py --gpus 8 style): Sponza
3840x2160, 16 spp, area + point + directional lights (add_point_light((-15,10,-5)*s, (20,20,20)),
add_directional_light(normalize(-0.3,-1,0.2), (3,3,3)), energy = |radiance|)."""
import json
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO / "rendering-fw_b200" / "python"))
import rfwb200 as R  # noqa: E402
import scenes as S  # noqa: E402

W, H, SPP = 3840, 2160, 16
sc = S.sponza_or_standin()
pl = np.zeros(1, R.POINT_LIGHT_DTYPE)
pl["position"], pl["radiance"] = (-15 * 0.2, 10 * 0.2, -5 * 0.2), (20, 20, 20)
pl["energy"] = np.linalg.norm(pl["radiance"][0])
dl = np.zeros(1, R.DIR_LIGHT_DTYPE)
d = np.array([-0.3, -1.0, 0.2])
dl["direction"], dl["radiance"] = d / np.linalg.norm(d), (3, 3, 3)
dl["energy"] = np.linalg.norm(dl["radiance"][0])
sc.point_lights, sc.dir_lights = pl, dl
ctx = R.RenderContext(R.load_product())
S.upload(ctx, sc, W, H)
ctx.set_setting("spp", SPP)
cam = sc.camera(W, H)
ms = []
for _ in range(4):
    ctx.render_frame(cam, R.RESET)
    ctx.synchronize()
    ms.append(ctx.get_stats().render_time)
img = ctx.read_image()
fc = ctx.get_frame_counters().as_dict()
t = float(np.median(ms[1:]))
print(json.dumps({"config": "sponza 3840x2160 16 spp, area + point + directional lights, 1 GPU", "frame_ms": t,
                  "msamples_per_s": W * H * SPP / (t * 1e3), "mean": float(img[..., :3].mean()), "finite": bool(np.isfinite(img).all()),
                  "counters": fc}))
