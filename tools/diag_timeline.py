#!/usr/bin/env python3
"""Per-warp timeline of one trace launch (start/end ns, rays, SM) to find stragglers."""
import ctypes as C, sys
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO / "rendering-fw_b200" / "python"))
import numpy as np
import rfwb200 as R, scenes as S
depth = int(sys.argv[1]) if len(sys.argv) > 1 else 2
sc = S.sponza_or_standin()
ctx = R.RenderContext(R.load_product())
S.upload(ctx, sc, 1920, 1080)
ctx.set_setting("spp", 1)
cam = sc.camera(1920, 1080)
ctx.render_frame(cam, R.RESET); ctx.synchronize()
f = ctx.L.fn("debug_trace_timeline", C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p])
n = C.c_size_t()
assert f(ctx._h, depth, None, 0, C.byref(n)) == 0
ctx.render_frame(cam, R.RESET); ctx.synchronize()
rec = np.zeros((n.value, 4), np.uint64)
assert f(ctx._h, depth, rec.ctypes.data, n.value, C.byref(n)) == 0
t0 = rec[:, 0].min()
start, end, rays, sm = (rec[:, 0] - t0) / 1e3, (rec[:, 1] - t0) / 1e3, rec[:, 2], rec[:, 3]
print("warps", len(rec), "rays total", rays.sum(), "start us: max %.1f" % start.max(), "end us: p50 %.1f p90 %.1f p99 %.1f max %.1f" % tuple(np.percentile(end, [50, 90, 99, 100])))
order = np.argsort(end)[::-1][:12]
for i in order:
    print("  warp %5d sm %3d start %.1f end %.1f rays %d" % (i, sm[i], start[i], end[i], rays[i]))
print("rays per warp: mean %.1f p1 %d p99 %d max %d" % (rays.mean(), np.percentile(rays, 1), np.percentile(rays, 99), rays.max()))
per_sm_end = {int(s): end[sm == s].max() for s in np.unique(sm)}
e = np.array(list(per_sm_end.values())); print("per-SM end us: min %.1f p50 %.1f max %.1f" % (e.min(), np.median(e), e.max()))
