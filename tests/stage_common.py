"""Helpers of the per-path shade parity tests (test infrastructure: the oracle entry used here, rfworacle_shade_stage,
has no counterpart in librfwb200.so — the product's shade stage is only reachable as part of a frame, and its planes are
read back through rfwb200_debug_read_plane).

Path buffers use the reference's layout (CUDART/src/Kernels.cu:571-592): O = (origin, bits((pathIndex << 8) | flags)),
D = (direction, bits(packed normal of the previous vertex)), T = (throughput, pdf), hit = (bits(w0_16 | w1_16 << 16),
bits(instance), bits(primitive) or -1, t)."""
import ctypes as C

import numpy as np

import rfwb200 as R

GEOMETRY_EPSILON = 1e-5


def f32x4(a):
    return np.ascontiguousarray(np.asarray(a, np.float32).reshape(-1, 4))


def bits(a):
    """float32 array -> the same bytes as uint32"""
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def as_float(a):
    return np.ascontiguousarray(a, np.uint32).view(np.float32)


def shade_stage(o, camera, O, D, T, hit, path_length, samples_taken=0):
    """rfworacle_shade_stage: shade_path of the oracle on every caller-supplied path.  Returns a dict of per-path arrays
    (unset outputs zeroed): flags (1 = extension, 2 = connect entry, 4 = accumulated), ext_O/D/T, con_O/D/E, acc."""
    O, D, T, hit = f32x4(O), f32x4(D), f32x4(T), f32x4(hit)
    n = len(O)
    assert len(D) == len(T) == len(hit) == n
    view = camera.get_view() if hasattr(camera, "get_view") else camera
    out = {k: np.zeros((n, 4), np.float32) for k in ("ext_O", "ext_D", "ext_T", "con_O", "con_D", "con_E", "acc")}
    flags = np.zeros(n, np.uint32)
    f = o.L.fn("shade_stage", C.c_int, [C.c_void_p, C.POINTER(R.CameraView)] + [C.c_void_p] * 4 + [C.c_size_t, C.c_uint32, C.c_uint32] + [C.c_void_p] * 8)
    o._check(f(o._h, C.byref(view), O.ctypes.data, D.ctypes.data, T.ctypes.data, hit.ctypes.data, n, path_length, samples_taken,
               flags.ctypes.data, *[out[k].ctypes.data for k in ("ext_O", "ext_D", "ext_T", "con_O", "con_D", "con_E", "acc")]))
    out["flags"] = flags
    return out


def hit_records(hits):
    """trace_closest's hits (Moller-Trumbore u, v) -> the path-state word the wavefront carries (oracle trace_to_state:
    weights of vertex 0 and vertex 1, 16 bits each, truncated)"""
    n = len(hits)
    rec = np.zeros((n, 4), np.uint32)
    miss = hits["prim_id"] < 0
    one = np.float32(1.0)
    w0 = (one - hits["u"].astype(np.float32)) - hits["v"].astype(np.float32)
    w1 = hits["u"].astype(np.float32)
    q = lambda w: np.clip(np.trunc(np.float32(65535.0) * w), 0, 4294967295).astype(np.uint32)
    rec[:, 0] = np.where(miss, 0, q(w0) | (q(w1) << np.uint32(16)))
    rec[:, 1] = np.where(miss, 0, hits["inst_id"]).astype(np.uint32)
    rec[:, 2] = np.where(miss, -1, hits["prim_id"]).astype(np.int32).view(np.uint32)
    rec[:, 3] = bits(np.where(miss, np.float32(0), hits["t"]).astype(np.float32))
    return as_float(rec).reshape(n, 4)


def compose_frame(o, camera, width, height, max_path_length):
    """One 1-sample PT frame put together from the oracle's stage entries in the order of its own host loop
    (render_sample_pt <- CUDART/src/Context.cpp:83-159): generate, extend, [shade, connect, extend]*.  Returns the
    accumulator (H, W, 4) and the per-depth queue sizes [(paths shaded, extension rays, connect entries)]."""
    P = width * height
    O, D = o.generate_primary(camera, 0)
    T = np.ones((P, 4), np.float32)
    hit = hit_records(o.trace_closest(O, D))
    acc = np.zeros((P, 4), np.float32)
    sizes = []
    depth = 0
    while True:
        s = shade_stage(o, camera, O, D, T, hit, depth, 0)
        pix = bits(O[:, 3]) >> 8
        a = (s["flags"] & 4) != 0
        acc[pix[a], :3] += s["acc"][a, :3]  # a path index occurs once per depth: no duplicate targets
        e, c = (s["flags"] & 1) != 0, (s["flags"] & 2) != 0
        sizes.append((len(O), int(e.sum()), int(c.sum())))
        if not (e.any() and depth < max_path_length):
            break
        depth += 1
        cO, cD, cE = s["con_O"][c], s["con_D"][c], s["con_E"][c]
        if len(cO):
            o4, d4 = cO.copy(), cD.copy()
            o4[:, 3] = 0
            d4[:, 3] = 0
            vis = o.trace_occluded(o4, d4, cD[:, 3].copy(), t_min=GEOMETRY_EPSILON) == 0
            tgt = bits(cE[:, 3])[vis]
            acc[tgt, :3] += cE[vis, :3]
            acc[tgt, 3] += 1.0
        O, D, T = s["ext_O"][e], s["ext_D"][e], s["ext_T"][e]
        o4, d4 = O.copy(), D.copy()
        o4[:, 3] = 0
        d4[:, 3] = 0
        hit = hit_records(o.trace_closest(o4, d4))
    return acc.reshape(height, width, 4), sizes


# ---- the product's shade stage per path (driven by tests/test_parity_scale_gpu.py on the GPU and, with an emulated product made
# of oracle stage calls, by tests/test_shade_stage.py on the CPU so that the joins below are themselves tested) ----------------
def read_counters(g, depths):
    cnt = np.zeros((8, 8), np.uint32)  # wavefront 0: one row per depth = (ext, shadow, trace_cursor, shade_cursor, acc, shadow_traced, -, -)
    g._check(g.L.fn("debug_read_counters", C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t])(g._h, cnt.ctypes.data, 8))
    return [(int(cnt[d, 0]), int(cnt[d, 1])) for d in range(depths)]


def _plane(g, which, n):
    return np.ascontiguousarray(g.debug_read_plane(which, n)).reshape(n, 4).copy()


def to_oracle_paths(O, hit, pix, shade):
    """planes of the product -> the oracle's conventions: path index = global pixel (the product carries the work item), hit record
    names (instance, primitive) instead of the flattened scene's shading record.  Returns O', hit', pixel per path, live mask."""
    w = bits(O[:, 3])
    pixel = pix[(w >> 8).astype(np.int64)]
    rec = bits(hit).reshape(-1, 4)
    z = rec[:, 2].view(np.int32)
    live = (z != -2) & (pixel >= 0)  # -2: padded work item of the tile grid
    miss = z == -1
    sid = np.minimum(rec[:, 1], len(shade) - 1).astype(np.int64)
    out = rec.copy()
    out[:, 1] = np.where(miss, 0, shade["inst_id"][sid]).astype(np.uint32)
    out[:, 2] = np.where(miss, -1, shade["prim_id"][sid].astype(np.int64)).astype(np.int32).view(np.uint32)
    O2 = O.copy()
    O2[:, 3] = as_float((np.maximum(pixel, 0).astype(np.uint32) << np.uint32(8)) | (w & np.uint32(255)))
    return O2, as_float(out).reshape(-1, 4), pixel, live


def _quantiles(e):
    if len(e) == 0:
        return "n/a"
    q = np.quantile(e, [0.5, 0.99, 0.999])
    return f"median {q[0]:.1e} p99 {q[1]:.1e} p99.9 {q[2]:.1e} max {e.max():.1e}"


def _compare_queue(tag, P, in_pixel, want_mask, want, got_pixel, got, tol, report):
    """`want[k]` = oracle outputs per INPUT path (rows where want_mask), `got` = the product's queue entries (any order, one per
    path); joined on the pixel.  Returns (fraction of the common entries whose every component is within tolerance, entries present on
    one side only as a fraction of the input paths)."""
    row_of = np.full(P, -1, np.int64)
    row_of[in_pixel] = np.arange(len(in_pixel))
    assert (got_pixel >= 0).all() and len(np.unique(got_pixel)) == len(got_pixel), f"{tag}: a path wrote two queue entries"
    rows = row_of[got_pixel]
    assert (rows >= 0).all(), f"{tag}: queue entry of a path that was not shaded"
    both = want_mask[rows]
    only_product = int((~both).sum())
    only_oracle = int(want_mask.sum() - both.sum())
    report[f"{tag}: entries"] = f"{len(got_pixel)} (oracle {int(want_mask.sum())}; only product {only_product}, only oracle {only_oracle})"
    one_sided = (only_product + only_oracle) / max(len(in_pixel), 1)
    r = rows[both]
    ok = np.ones(len(r), bool)
    for k, (ref, val, kind) in got.items():
        a, b = val[both], want[ref][r]
        if kind == "pos":  # points of the scene: relative to the coordinate magnitude
            e = np.abs(a[:, :3] - b[:, :3]).max(1) / np.maximum(1.0, np.abs(b[:, :3]).max(1))
            lim = 1e-5
        elif kind == "dir":
            e = np.abs(a[:, :3] - b[:, :3]).max(1)
            lim = tol
        elif kind == "len":
            e = np.abs(a[:, 3] - b[:, 3]) / np.maximum(1.0, np.abs(b[:, 3]))
            lim = 1e-5
        else:  # radiometric values (throughput, pdf, contribution)
            cols = slice(0, 4) if kind == "val4" else slice(0, 3)
            e = (np.abs(a[:, cols] - b[:, cols]) / (np.abs(b[:, cols]) + 1e-2)).max(1)
            lim = tol
        e = np.where(np.isfinite(e), e, np.inf)
        report[f"{tag}: {k}"] = _quantiles(e) + f"; within {lim:.0e}: {(e <= lim).mean():.5f}"
        ok &= e <= lim
    return (float(ok.mean()) if len(ok) else 1.0), one_sided


def check_shade_stage_per_path(g, o, cam, W, H, tol, min_queue):
    """Renders three 1-sample frames on the product `g` (path length 0, 1, 2), reads the planes its shade launches read and wrote,
    runs the oracle's shade_path on the same (ray, hit) pairs and compares every output per path.  Returns ({what: fraction of the
    paths within tolerance — keys ending in "one-sided": fraction of the paths with an entry on one side only}, {what: description}).  `o` must hold the same scene with max_path_length >= 2."""
    P = W * H
    pix = np.asarray(R.shard_pixel_map(W, H, 0, 1))
    n_items = len(pix)
    shade = g.debug_read_scene("shade")
    report, fr = {}, {}

    # frame A (path length 0 only): the camera rays, their hits, and what shade(0) accumulated
    g.set_setting("max_path_length", 0)
    g.render_frame(cam, R.RESET)
    O0, D0, hit0, acc0 = _plane(g, 0, n_items), _plane(g, 2, n_items), _plane(g, 6, n_items), _plane(g, 10, n_items)
    # frame B (one bounce): outputs of shade(0) = the re-ordered depth-1 queue and the connect queue; the depth-1 hits; per-sample radiance
    g.set_setting("max_path_length", 1)
    g.render_frame(cam, R.RESET)
    (n_ext0, n_con0), = read_counters(g, 1)
    assert n_ext0 >= min_queue and n_con0 >= min_queue, (n_ext0, n_con0)
    E_O, E_D, E_T, hit1 = _plane(g, 0, n_ext0), _plane(g, 2, n_ext0), _plane(g, 4, n_ext0), _plane(g, 6, n_ext0)
    C_O, C_D, C_E = _plane(g, 7, n_con0), _plane(g, 8, n_con0), _plane(g, 9, n_con0)
    accB = _plane(g, 10, n_items)
    # frame C (two bounces): outputs of shade(1)
    g.set_setting("max_path_length", 2)
    g.render_frame(cam, R.RESET)
    (n_ext0c, n_con0c), (n_ext1, n_con1) = read_counters(g, 2)
    assert (n_ext0c, n_con0c) == (n_ext0, n_con0)  # shade(0) does not depend on how long paths may become
    F_O, F_D, F_T = _plane(g, 0, n_ext1), _plane(g, 2, n_ext1), _plane(g, 4, n_ext1)
    G_O, G_D, G_E = _plane(g, 7, n_con1), _plane(g, 8, n_con1), _plane(g, 9, n_con1)

    item_pixel = lambda word: pix[word.astype(np.int64)]
    ext_spec = lambda o_, d_, t_: {"origin": ("ext_O", o_, "pos"), "direction": ("ext_D", d_, "dir"), "throughput, pdf": ("ext_T", t_, "val4")}
    con_spec = lambda o_, d_, e_: {"origin": ("con_O", o_, "pos"), "direction": ("con_D", d_, "dir"), "length": ("con_D", d_, "len"),
                                   "contribution": ("con_E", e_, "val3")}

    # ---- shade(0) ----
    O0o, hit0o, pixel0, live0 = to_oracle_paths(O0, hit0, pix, shade)
    assert live0.sum() == P and np.array_equal(np.sort(pixel0[live0]), np.arange(P))
    s0 = shade_stage(o, cam, O0o[live0], D0[live0], np.ones_like(O0[live0]), hit0o[live0], 0, 0)
    in_pixel0 = pixel0[live0]
    fr["shade(0) extension"], fr["shade(0) extension: one-sided"] = _compare_queue("shade(0) extension", P, in_pixel0, (s0["flags"] & 1) != 0, s0, item_pixel(bits(E_O[:, 3]) >> 8),
                                              ext_spec(E_O, E_D, E_T), tol, report)
    fr["shade(0) connect"], fr["shade(0) connect: one-sided"] = _compare_queue("shade(0) connect", P, in_pixel0, (s0["flags"] & 2) != 0, s0, item_pixel(bits(C_E[:, 3])),
                                            con_spec(C_O, C_D, C_E), tol, report)
    # flags of the extension rays (specular bit) and the packed normal they carry
    rows = np.full(P, -1, np.int64)
    rows[in_pixel0] = np.arange(P)
    r = rows[item_pixel(bits(E_O[:, 3]) >> 8)]
    common = (s0["flags"][r] & 1) != 0
    same_flags = (bits(E_O[:, 3])[common] & 255) == (bits(s0["ext_O"][r[common], 3]) & 255)
    same_normal = bits(E_D[:, 3])[common] == bits(s0["ext_D"][r[common], 3])
    report["shade(0) extension: flag byte equal"] = f"{same_flags.mean():.6f}; packed normal word equal: {same_normal.mean():.5f}"
    fr["shade(0) flag byte"] = float(same_flags.mean())
    # accumulated at depth 0 (sky, emitters seen directly): every pixel
    a_ref = np.zeros((P, 3), np.float32)
    a_ref[in_pixel0] = s0["acc"][:, :3]
    a_got = np.zeros((P, 3), np.float32)
    a_got[in_pixel0] = acc0[live0, :3]
    e = (np.abs(a_got - a_ref) / (np.abs(a_ref) + 1e-2)).max(1)
    fr["shade(0) accumulated"] = float((e <= tol).mean())
    report["shade(0) accumulated"] = _quantiles(e) + f"; paths that accumulate: {int(((s0['flags'] & 4) != 0).sum())}"

    # ---- shade(1): the rays of the depth-1 launch with the hits it found ----
    E_Oo, hit1o, pixel1, live1 = to_oracle_paths(E_O, hit1, pix, shade)
    assert live1.all()
    s1 = shade_stage(o, cam, E_Oo, E_D, E_T, hit1o, 1, 0)
    fr["shade(1) extension"], fr["shade(1) extension: one-sided"] = _compare_queue("shade(1) extension", P, pixel1, (s1["flags"] & 1) != 0, s1, item_pixel(bits(F_O[:, 3]) >> 8),
                                              ext_spec(F_O, F_D, F_T), tol, report)
    fr["shade(1) connect"], fr["shade(1) connect: one-sided"] = _compare_queue("shade(1) connect", P, pixel1, (s1["flags"] & 2) != 0, s1, item_pixel(bits(G_E[:, 3])),
                                            con_spec(G_O, G_D, G_E), tol, report)

    # ---- the one-bounce frame per pixel: oracle outputs + the product's own visibility decisions (the sample slot's .w counts the
    # connect contributions that arrived: 0 or 1 with one bounce) ----
    total = a_ref.copy()
    c0 = (s0["flags"] & 2) != 0
    arrived = np.zeros(P, bool)
    arrived[in_pixel0] = accB[live0, 3] == 1.0
    con = np.zeros((P, 3), np.float32)
    emitted = np.zeros(P, bool)
    con[in_pixel0[c0]] = s0["con_E"][c0, :3]
    emitted[in_pixel0[c0]] = True
    total += np.where(arrived[:, None], con, 0)
    a1 = (s1["flags"] & 4) != 0
    total[pixel1[a1]] += s1["acc"][a1, :3]
    got = np.zeros((P, 3), np.float32)
    got[in_pixel0] = accB[live0, :3]
    e = (np.abs(got - total) / (np.abs(total) + 1e-2)).max(1)
    fr["frame radiance (one bounce)"] = float((e <= tol).mean())
    report["frame radiance (one bounce)"] = (_quantiles(e) + f"; connect contributions that arrived: {int(arrived.sum())} of {int(c0.sum())}; "
                                             f"arrived without an oracle entry: {int((arrived & ~emitted).sum())}")
    return fr, report
