#!/usr/bin/env python3
"""A/B of the traversal-kernel instantiations on BASELINE.json configs[1] (Sponza 1920x1080 8 spp): frame time (CUDA events)
and per-stage times (`timing`) per setting.  One JSON line per variant."""
import argparse
import json
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO / "rendering-fw_b200" / "python"))
import rfwb200 as R  # noqa: E402
import scenes as S  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=12)
ap.add_argument("--spp", type=int, default=8)
ap.add_argument("--width", type=int, default=1920)
ap.add_argument("--height", type=int, default=1080)
ap.add_argument("--variants", default="default;trace_variant=0;trace_variant=5;trace_variant=8;trace_variant=9;trace_variant=10;default",
                help="';'-separated variants, each a ','-separated list of key=value settings ('default' = none)")
a = ap.parse_args()

sc = S.sponza_or_standin()
ctx = R.RenderContext(R.load_product())
S.upload(ctx, sc, a.width, a.height)
ctx.set_setting("spp", a.spp)
cam = sc.camera(a.width, a.height)
ref = None
for var in a.variants.split(";"):
    kvs = [kv.split("=") for kv in var.split(",") if kv and kv != "default"]
    for k, v in kvs:
        ctx.set_setting(k, v)
    ctx.update()
    ctx.set_setting("timing", "off")
    ms = []
    for _ in range(a.frames):
        ctx.render_frame(cam, R.RESET)
        ctx.synchronize()
        ms.append(ctx.get_stats().render_time)
    img = ctx.read_image()
    if ref is None:
        ref = img.copy()
    ctx.set_setting("timing", "on")
    ctx.render_frame(cam, R.RESET)
    ctx.synchronize()
    st = ctx.get_stats()
    print(json.dumps({"variant": var, "frame_ms_median": float(np.median(ms[2:])), "frame_ms_min": float(np.min(ms[2:])),
                      "msamples_per_s": a.width * a.height * a.spp / (float(np.median(ms[2:])) * 1e3),
                      "stages": {"primary": st.primary_time, "trace_d1": st.secondary_time, "trace_d2": st.deep_time, "shade": st.shade_time,
                                 "sort": st.animation_time, "fold": st.finalize_time},
                      "pixels_differing_from_first": float((np.abs(img - ref).max(axis=-1) > 0).mean()),
                      "mean": float(img[..., :3].mean()), "bvh": ctx.get_bvh_info()}), flush=True)
    for k, v in kvs:  # back to the defaults
        ctx.set_setting(k, {"trace_variant": "9", "primary_variant": "9", "bvh": "4", "shadow_cache": "off", "primary_cache": "on", "sort": "on",
                               "sort_cell_bits": "5", "sort_major": "cell", "spp_batch": "0", "fetch_threshold": "8", "fetch_chunk": "0", "sample_layout": "pixel", "sort_dir_bits": "5", "shade_loop": "cursor"}.get(k, v))
    ctx.set_setting("primary_variant", "9")
