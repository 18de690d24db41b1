#!/usr/bin/env python3
"""Diagnose straggler rays: dump the depth-2 extension / connect queues of one Sponza sample and time sub-batches."""
import ctypes as C, sys, time
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO / "rendering-fw_b200" / "python"))
import numpy as np
import rfwb200 as R, scenes as S
sc = S.sponza_or_standin()
ctx = R.RenderContext(R.load_product())
W, H = 1920, 1080
S.upload(ctx, sc, W, H)
ctx.set_setting("spp", 1)
ctx.render_frame(sc.camera(W, H), R.RESET)
ctx.synchronize()
cnt = np.zeros((8, 8), np.uint32)
ctx.L.fn("debug_read_counters", C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t])(ctx._h, cnt.ctypes.data, 8)
print("counters ext/shadow per depth", cnt[:3, :2].tolist())
def plane(which, n):
    a = np.zeros((n, 4), np.float32)
    rc = ctx.L.fn("debug_read_plane", C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t])(ctx._h, which, a.ctypes.data, n)
    assert rc == 0
    return a
n_ext2, n_sh2 = int(cnt[1, 0]), int(cnt[1, 1])
O, D = plane(0, n_ext2), plane(2, n_ext2)      # depth-2 extension rays live in buffer 0
sO, sD = plane(7, n_sh2), plane(8, n_sh2)      # connect rays emitted by shade(1)
for name, o, d in (("ext2", O, D), ("shadow2", sO, sD)):
    print(name, len(o), "nan", np.isnan(o[:, :3]).any(axis=1).sum(), np.isnan(d[:, :3]).any(axis=1).sum(),
          "zero-dir", (np.abs(d[:, :3]).sum(axis=1) == 0).sum(), "|d| range", np.linalg.norm(d[:, :3], axis=1).min(), np.linalg.norm(d[:, :3], axis=1).max(),
          "|o| max", np.abs(o[:, :3]).max(), "tmax range" if name == "shadow2" else "", (d[:, 3].min(), d[:, 3].max()) if name == "shadow2" else "")
# time chunks of the extension rays through the stage-level entry point
chunk = 1 << 16
ts = []
for i in range(0, len(O), chunk):
    t0 = time.perf_counter(); ctx.trace_closest(O[i:i + chunk], D[i:i + chunk]); ts.append(time.perf_counter() - t0)
ts = np.array(ts); print("ext2 chunk ms: median %.2f max %.2f argmax %d of %d" % (np.median(ts) * 1e3, ts.max() * 1e3, ts.argmax(), len(ts)))
ts = []
for i in range(0, len(sO), chunk):
    t0 = time.perf_counter(); ctx.trace_occluded(sO[i:i + chunk], sD[i:i + chunk], sD[i:i + chunk, 3]); ts.append(time.perf_counter() - t0)
ts = np.array(ts); print("shadow2 chunk ms: median %.2f max %.2f argmax %d of %d" % (np.median(ts) * 1e3, ts.max() * 1e3, ts.argmax(), len(ts)))
np.savez_compressed(REPO / "gpurun_out" / "rays_d2.npz", O=O[::50], D=D[::50], sO=sO[::50], sD=sD[::50])
