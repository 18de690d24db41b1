#!/usr/bin/env python3
"""In-process device group (rfwb200_create_group, what the B200RT plugin uses with RFWB200_DEVICES): ONE context, ONE host
thread calling render_frame, N GPUs.  Renders a bench config, checks the frame against a single-device render of it bit for
bit, and times it (render_frame + the display wait, CUDA events of rank 0).  One JSON line."""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO / "rendering-fw_b200" / "python"))
sys.path.insert(0, str(REPO))
import bench as B  # noqa: E402
import rfwb200 as R  # noqa: E402
import scenes as S  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--devices", default="0,1")
ap.add_argument("--config", type=int, default=2)
ap.add_argument("--frames", type=int, default=30)
a = ap.parse_args()
devices = [int(x) for x in a.devices.split(",")]
cfg = B.CONFIGS[a.config]
W, H, SPP = cfg["width"], cfg["height"], cfg["spp"]
lib = R.load_product()
sc, skins = B.build_scene(cfg)
cam = sc.camera(W, H)


def make(devs):
    ctx = R.RenderContext(lib, devices=devs) if len(devs) > 1 else R.RenderContext(lib, devs[0])
    t0 = time.perf_counter()
    S.upload(ctx, sc, W, H)
    up = time.perf_counter() - t0
    ctx.set_setting("spp", SPP), ctx.set_setting("max_path_length", B.MAX_PATH)
    return ctx, up


grp, up_g = make(devices)
for _ in range(3):
    grp.render_frame(cam, R.RESET)
grp.synchronize()
ms = []
t0 = time.perf_counter()
for _ in range(a.frames):
    grp.render_frame(cam, R.RESET)
    grp.synchronize()
    ms.append(grp.get_stats().render_time)
wall = (time.perf_counter() - t0) * 1e3 / a.frames
t0 = time.perf_counter()
for _ in range(a.frames):  # back to back, one synchronisation at the end: what an application that keeps rendering sees
    grp.render_frame(cam, R.RESET)
grp.synchronize()
pipelined = (time.perf_counter() - t0) * 1e3 / a.frames
img = grp.read_image().copy()
one, up_1 = make(devices[:1])
one.render_frame(cam, R.RESET)
ref = one.read_image()
t = float(np.median(ms))
print(json.dumps({"config": a.config, "devices": devices, "frame_ms_events": t, "frame_ms_wall_sync_each": wall, "frame_ms_wall_back_to_back": pipelined,
                  "msamples_per_s": W * H * SPP / (pipelined * 1e3), "identical_to_one_device": bool(np.array_equal(img, ref)),
                  "pixels_differing": int((np.abs(img - ref).max(axis=-1) > 0).sum()), "mean": float(img[..., :3].mean()),
                  "upload_s": {"group": up_g, "one": up_1}, "counters_equal": grp.get_frame_counters().as_dict() == one.get_frame_counters().as_dict()}))
