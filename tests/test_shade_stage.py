"""The oracle's stage-level shade entry (rfworacle_shade_stage, used by the per-path GPU parity tests at the benchmarked
scale) is checked first: a frame put together from generate + extend + shade + connect stage calls, in the order of the
oracle's own host loop, must equal the oracle's frame BIT FOR BIT — the oracle's frames are what is pinned on the
reference's own kernels (tests/test_ref_pin.py), so the stage entry inherits that pin."""
import numpy as np
import pytest

import rfwb200 as R
import scenes as S
from stage_common import compose_frame, hit_records, shade_stage


def unit_cornell():
    return S.cornell_box(unit_scale=True)


@pytest.mark.parametrize("scene", ["cornell", "soup"])
@pytest.mark.parametrize("depth", [0, 1, 2])
def test_frame_composed_from_stage_calls_equals_the_oracle_frame(oracle_lib, scene, depth):
    W, H = 96, 64
    sc = {"cornell": unit_cornell, "soup": S.feature_soup}[scene]()
    o = R.RenderContext(oracle_lib)
    S.upload(o, sc, W, H)
    o.set_setting("spp", 1)
    o.set_setting("max_path_length", depth)
    cam = sc.camera(W, H)
    o.render_frame(cam, R.RESET)
    frame = o.read_image().copy()
    counters = o.get_frame_counters().as_dict()
    acc, sizes = compose_frame(o, cam, W, H, depth)
    assert np.array_equal(acc[..., :3], frame[..., :3])
    assert sum(s[0] for s in sizes) == counters["n_shade"]
    assert sum(s[1] for s in sizes) == counters["n_ext_out"]
    assert sizes[0][0] == W * H and len(sizes) <= depth + 1


def test_stage_outputs_are_zeroed_where_unset_and_reject_bad_records(oracle_lib):
    W, H = 64, 48
    sc = unit_cornell()
    o = R.RenderContext(oracle_lib)
    S.upload(o, sc, W, H)
    o.set_setting("max_path_length", 2)
    cam = sc.camera(W, H)
    O, D = o.generate_primary(cam, 0)
    hit = hit_records(o.trace_closest(O, D))
    s = shade_stage(o, cam, O, D, np.ones_like(O), hit, 0, 0)
    e, c, a = (s["flags"] & 1) != 0, (s["flags"] & 2) != 0, (s["flags"] & 4) != 0
    assert e.any() and c.any() and a.any()  # the box shows its light, diffuse walls and NEE
    assert not (a & (e | c)).any()          # a path that accumulates (sky, emitter) ends
    for k, m in (("ext_O", e), ("ext_D", e), ("ext_T", e), ("con_O", c), ("con_D", c), ("con_E", c), ("acc", a)):
        assert not s[k][~m].any(), k
    # extension rays keep their path index and carry a finite throughput and a positive pdf
    assert np.array_equal(s["ext_O"][e, 3].view(np.uint32) >> 8, np.nonzero(e)[0].astype(np.uint32))
    assert np.isfinite(s["ext_T"][e]).all() and (s["ext_T"][e, 3] >= 1e-6).all()
    # a connect entry names its path and stops short of the light
    assert np.array_equal(s["con_E"][c, 3].view(np.uint32), np.nonzero(c)[0].astype(np.uint32))
    assert (s["con_D"][c, 3] > 0).all()
    bad = hit.copy()
    bad.view(np.uint32)[0, 1] = 10_000  # unknown instance
    bad.view(np.uint32)[0, 2] = 0
    with pytest.raises(R.Rfwb200Error):
        shade_stage(o, cam, O, D, np.ones_like(O), bad, 0, 0)


class EmulatedProduct:
    """Stands in for the product in a CPU run of the per-path checker (stage_common.check_shade_stage_per_path): frames made of oracle
    stage calls, kept in planes the way librfwb200's frame keeps them — work items instead of pixels in the path word, shading-record
    indices instead of (instance, primitive) in the hit word, queues in another order than they were emitted, per-depth counters,
    per-sample radiance with the arrival count in .w.  `spoil` scales the throughput plane of the depth-1 queue (a wrong product)."""

    def __init__(self, o, sc, W, H, spoil=1.0):
        self.o, self.W, self.H, self.spoil = o, W, H, spoil
        self.pix = np.asarray(R.shard_pixel_map(W, H, 0, 1))
        self.n_items = len(self.pix)
        self.item_of_pixel = np.full(W * H, -1, np.int64)
        self.item_of_pixel[self.pix[self.pix >= 0]] = np.nonzero(self.pix >= 0)[0]
        counts = [len(sc.meshes[m].triangles) for m, _ in sc.instances]
        self.offset = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        n = int(self.offset[-1])
        rng = np.random.default_rng(5)
        self.where = rng.permutation(n)  # (instance, primitive) -> shading record
        self.shade = np.zeros(n, np.dtype([("inst_id", "<u4"), ("prim_id", "<u4")]))
        inst = np.repeat(np.arange(len(counts)), counts)
        self.shade["inst_id"][self.where] = inst
        self.shade["prim_id"][self.where] = np.arange(n) - self.offset[inst]
        self.planes = {k: np.zeros((self.n_items, 4), np.float32) for k in range(11)}
        self.counters = np.zeros((8, 8), np.uint32)
        self.mpl, self.rng, self.L, self._h = 2, rng, self, None

    # -- the slice of the RenderContext / Library interface the checker uses --
    def set_setting(self, k, v):
        assert k == "max_path_length"
        self.mpl = int(v)

    def debug_read_scene(self, which):
        assert which == "shade"
        return self.shade

    def debug_read_plane(self, which, n):
        return self.planes[which][:n].reshape(-1).copy()

    def _check(self, rc):
        assert rc == 0

    def fn(self, name, restype=None, argtypes=None):
        import ctypes as C
        assert name == "debug_read_counters"
        return lambda h, ptr, slots: C.memmove(ptr, self.counters.ctypes.data, slots * 32) and 0

    # -- a frame --
    def _product_hits(self, rec):
        """oracle path-state hit words -> the product's (record index in .y, 0 / -1 in .z)"""
        from stage_common import as_float, bits
        r = bits(rec).reshape(-1, 4).copy()
        prim = r[:, 2].view(np.int32)
        miss = prim < 0
        idx = self.where[np.where(miss, 0, self.offset[np.minimum(r[:, 1], len(self.offset) - 2)] + np.maximum(prim, 0))]
        r[:, 1] = np.where(miss, 0, idx).astype(np.uint32)
        r[:, 2] = np.where(miss, -1, 0).astype(np.int32).view(np.uint32)
        return as_float(r).reshape(-1, 4)

    def _items(self, word, shift):
        """path words (pixel << shift | low bits) -> (item << shift | low bits)"""
        from stage_common import as_float, bits
        w = bits(word)
        low = w & np.uint32((1 << shift) - 1)
        return as_float((self.item_of_pixel[(w >> np.uint32(shift)).astype(np.int64)].astype(np.uint32) << np.uint32(shift)) | low)

    def render_frame(self, cam, status):
        from stage_common import bits, hit_records, shade_stage
        o, pl = self.o, self.planes
        for p in pl.values():
            p[:] = 0
        self.counters[:] = 0
        O, D = o.generate_primary(cam, 0)
        T = np.ones_like(O)
        hit = hit_records(o.trace_closest(O, D))
        it = self.item_of_pixel
        pl[0][it, :3], pl[0][it, 3] = O[:, :3], self._items(O[:, 3], 8)
        pl[2][it] = D
        dead = np.zeros((self.n_items, 4), np.int32)
        dead[:, 2] = -2
        pl[6][:] = dead.view(np.float32)
        pl[6][it] = self._product_hits(hit)
        item_of_path = it.copy()  # per row of the current wavefront
        for depth in range(self.mpl + 1):
            s = shade_stage(o, cam, O, D, T, hit, depth, 0)
            a = (s["flags"] & 4) != 0
            pl[10][item_of_path[a], :3] += s["acc"][a, :3]
            if depth >= self.mpl:
                break
            e, c = np.nonzero(s["flags"] & 1)[0], np.nonzero(s["flags"] & 2)[0]
            e, c = self.rng.permutation(e), self.rng.permutation(c)  # queue order is not emission order
            self.counters[depth, 0], self.counters[depth, 1] = len(e), len(c)
            cO, cD, cE = s["con_O"][c], s["con_D"][c], s["con_E"][c]
            pl[7][: len(c)], pl[8][: len(c)] = cO, cD
            pl[9][: len(c), :3], pl[9][: len(c), 3] = cE[:, :3], self._items(cE[:, 3], 0)
            if not len(e):
                break
            vis = o.trace_occluded(cO, cD, cD[:, 3].copy(), t_min=1e-5) == 0
            tgt = it[bits(cE[:, 3])[vis].astype(np.int64)]
            pl[10][tgt, :3] += cE[vis, :3]
            pl[10][tgt, 3] += 1.0
            O, D, T = s["ext_O"][e], s["ext_D"][e], s["ext_T"][e]
            hit = hit_records(o.trace_closest(O, D))
            item_of_path = it[(bits(O[:, 3]) >> 8).astype(np.int64)]
            n = len(e)
            pl[0][:n, :3], pl[0][:n, 3] = O[:, :3], self._items(O[:, 3], 8)
            pl[2][:n], pl[4][:n] = D, T * np.float32(self.spoil if depth == 0 else 1.0)
            pl[6][:n] = self._product_hits(hit)


@pytest.mark.parametrize("size", [(96, 64), (70, 50)])  # the second leaves padded work items in the tile grid
def test_per_path_checker_on_an_emulated_product(oracle_lib, size):
    """The checker the GPU test runs at 1920x1080, run here against a product emulated with the oracle's own stage calls (planes,
    work-item path words, shading-record hit words, shuffled queues): everything must agree exactly, and a product whose depth-1
    throughputs are 1 % off must be caught."""
    from stage_common import check_shade_stage_per_path
    W, H = size
    sc = S.feature_soup()
    o = R.RenderContext(oracle_lib)
    S.upload(o, sc, W, H)
    o.set_setting("spp", 1)
    o.set_setting("max_path_length", 2)
    cam = sc.camera(W, H)
    fr, report = check_shade_stage_per_path(EmulatedProduct(o, sc, W, H), o, cam, W, H, 1e-6, 100)
    assert all(v == (0.0 if k.endswith("one-sided") else 1.0) for k, v in fr.items()), (fr, report)
    assert "only product 0, only oracle 0" in report["shade(0) extension: entries"] and "only product 0, only oracle 0" in report["shade(1) connect: entries"]
    fr, report = check_shade_stage_per_path(EmulatedProduct(o, sc, W, H, spoil=1.01), o, cam, W, H, 1e-3, 100)
    # (shade(1) is then given the spoiled throughputs on both sides, so only the stage that wrote them stands out)
    assert fr["shade(0) extension"] < 0.5 and fr["shade(0) connect"] == 1.0 and fr["shade(0) accumulated"] == 1.0, fr


def test_texel_fetch_is_decided_by_the_last_bit_in_the_reference_formula(oracle_lib):
    """Why ~0.1 % of the paths of the per-path GPU check (tests/test_parity_scale_gpu.py) carry another COLOUR while every origin,
    direction, pdf and flag agrees: the reference's FetchTexel (CUDART/src/getShadingData.h:29-59) evaluates
    (tc + 1000) * width - 0.5 in float32 — at 1000 one ulp is 6e-5, i.e. 1/8 of a texel of a 2048-wide map — so the last bit of the
    interpolated texture coordinate moves the bilinear footprint by a fraction of a texel, and the texel enters the colour squared
    (:160,213).  Shown here without any GPU: the oracle built with FMA contraction (librfworacle_fast.so; nvcc contracts too) against
    the strict build, on identical rays and hit records of the headline scene — same entries, same geometry, and the same class of
    colour outliers (measured at 960x540: 0.15 % of the connect contributions beyond 1e-4, worst 0.3)."""
    from oracle.oracle_lib import load_oracle

    W, H = 480, 270
    ctxs = []
    for lib in (oracle_lib, load_oracle(fast=True)):
        sc = S.sponza_or_standin()
        o = R.RenderContext(lib)
        S.upload(o, sc, W, H)
        o.set_setting("spp", 1)
        o.set_setting("max_path_length", 2)
        ctxs.append(o)
    cam = sc.camera(W, H)
    O, D = ctxs[0].generate_primary(cam, 0)
    hit = hit_records(ctxs[0].trace_closest(O, D))
    strict, fma = (shade_stage(c, cam, O, D, np.ones_like(O), hit, 0, 0) for c in ctxs)
    assert np.array_equal(strict["flags"], fma["flags"])
    e, c = (strict["flags"] & 1) != 0, (strict["flags"] & 2) != 0
    assert e.sum() > 0.3 * W * H and c.sum() > 0.5 * W * H
    rel = lambda k, m, cols: (np.abs(fma[k][m, :cols] - strict[k][m, :cols]) / (np.abs(strict[k][m, :cols]) + 1e-2)).max(1)
    for k, m, cols in (("ext_O", e, 3), ("ext_D", e, 3), ("con_O", c, 3), ("con_D", c, 4)):
        assert rel(k, m, cols).max() <= 2e-4, (k, rel(k, m, cols).max())
    pdf = np.abs(fma["ext_T"][e, 3] - strict["ext_T"][e, 3]) / (np.abs(strict["ext_T"][e, 3]) + 1e-2)
    colour = rel("con_E", c, 3)
    print(f"FMA contraction on identical inputs: pdf max {pdf.max():.1e}; connect contributions beyond 1e-4: {(colour > 1e-4).mean():.5f}, worst {colour.max():.2f}")
    assert (colour > 1e-4).mean() <= 5e-3 and np.median(colour) <= 1e-5
    if any(t["width"] >= 1024 for t in sc.textures):
        # the textured asset: the outliers exist (a procedural stand-in without large textures need not show them)
        assert (colour > 1e-3).sum() >= 1
