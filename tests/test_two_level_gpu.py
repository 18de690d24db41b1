"""Two-level scenes (setting "levels" = 2 | auto): a top-level tree over instances + one object-space tree per mesh, the
shape the reference itself traverses (CUDART/src/Kernels.cu:226-303, CUDAIntersect.h:270-322; top level built as in
RFW/system/bvh/src/top_level_bvh.cpp:17-102).  The CPU oracle is two-level as well, so these compare like with like:
per-ray hits, E-mode and PT images, counters, probe; plus the flattened frame of the same scene, an instance that
moves, and the automatic switch on the flatten budget.  Tolerances are those of tests/test_parity_gpu.py."""
import numpy as np
import pytest

import rfwb200 as R
import scenes as S
from test_parity_gpu import SCENES, _check_hits, frac_bad, make_pair

pytestmark = pytest.mark.gpu


def product(product_lib, scene_fn, W, H, **settings):
    sc = scene_fn()
    ctx = R.RenderContext(product_lib)
    for k, v in settings.items():
        ctx.set_setting(k, v)
    S.upload(ctx, sc, W, H)
    return ctx, sc


def oracle(oracle_lib, scene_fn, W, H, **settings):
    sc = scene_fn()
    ctx = R.RenderContext(oracle_lib)
    S.upload(ctx, sc, W, H)
    for k, v in settings.items():
        ctx.set_setting(k, v)
    return ctx


@pytest.mark.parametrize("scene", ["cornell", "soup"])
def test_two_level_hits_match_the_oracle(product_lib, oracle_lib, scene):
    W, H = 160, 120
    g, sc = product(product_lib, SCENES[scene], W, H, levels=2)
    assert "levels_in_use=2" in g.get_settings()
    o = oracle(oracle_lib, SCENES[scene], W, H)
    origins, dirs = o.generate_primary(sc.camera(W, H), 0)
    hg, ho = g.trace_closest(origins, dirs), o.trace_closest(origins, dirs)
    assert (ho["prim_id"] >= 0).mean() > 0.3
    _check_hits(g, o, origins, dirs, hg, ho)
    # incoherent rays from inside the scene, and their occlusion answers
    rng = np.random.default_rng(5)
    hitp = origins[:, :3] + dirs[:, :3] * np.minimum(ho["t"], 1e3)[:, None] * 0.5
    sel = rng.choice(len(hitp), 4000, replace=False)
    o2 = np.zeros((4000, 4), np.float32)
    d2 = np.zeros((4000, 4), np.float32)
    o2[:, :3] = hitp[sel]
    v = rng.normal(size=(4000, 3))
    d2[:, :3] = v / np.linalg.norm(v, axis=1, keepdims=True)
    hg2, ho2 = g.trace_closest(o2, d2), o.trace_closest(o2, d2)
    _check_hits(g, o, o2, d2, hg2, ho2)
    tmax = np.where(ho2["prim_id"] >= 0, ho2["t"] * 1.5, 10.0).astype(np.float32)
    og, oo = g.trace_occluded(o2, d2, tmax), o.trace_occluded(o2, d2, tmax)
    assert (og != oo).mean() < 1e-3


@pytest.mark.parametrize("scene", ["cornell", "soup"])
def test_two_level_emode_image_and_probe(product_lib, oracle_lib, scene):
    W, H = 256, 192
    g, sc = product(product_lib, SCENES[scene], W, H, levels=2, mode="embree")
    o = oracle(oracle_lib, SCENES[scene], W, H, mode="embree")
    cam = sc.camera(W, H)
    g.set_probe_index(W // 2, H // 2), o.set_probe_index(W // 2, H // 2)
    g.render_frame(cam, R.RESET), o.render_frame(cam, R.RESET)
    a, b = g.read_image(), o.read_image()
    assert np.isfinite(a).all()
    assert frac_bad(a, b) < (1e-3 if scene == "cornell" else 5e-3)
    pg, po = g.get_probe_results(), o.get_probe_results()
    assert pg[:2] == po[:2] and abs(pg[2] - po[2]) <= 1e-4 * max(1.0, po[2])


@pytest.mark.parametrize("scene", ["cornell", "soup"])
@pytest.mark.parametrize("depth", [0, 2])
def test_two_level_pt_image_and_counters(product_lib, oracle_lib, scene, depth):
    W, H = 192, 128
    g, sc = product(product_lib, SCENES[scene], W, H, levels=2, max_path_length=depth, spp=1)
    o = oracle(oracle_lib, SCENES[scene], W, H, max_path_length=depth, spp=1)
    cam = sc.camera(W, H)
    g.set_probe_index(W // 3, H // 2), o.set_probe_index(W // 3, H // 2)
    g.render_frame(cam, R.RESET), o.render_frame(cam, R.RESET)
    a, b = g.read_image(), o.read_image()
    assert np.isfinite(a).all()
    limit = {0: 2e-3, 2: 8e-3}[depth] * (1 if scene == "cornell" else 3)
    assert frac_bad(a, b) < limit, (frac_bad(a, b), limit)
    cg, co = g.get_frame_counters().as_dict(), o.get_frame_counters().as_dict()
    assert cg["n_gen"] == co["n_gen"] and cg["pixels"] == co["pixels"]
    for k in ("n_ext", "n_shade", "n_ext_out", "n_nee", "n_acc"):
        assert abs(cg[k] - co[k]) <= 2e-3 * max(co[k], 1) + 2, (k, cg[k], co[k])
    pg, po = g.get_probe_results(), o.get_probe_results()
    assert pg[:2] == po[:2], (pg, po)


@pytest.mark.parametrize("scene", ["cornell", "soup"])
def test_two_level_and_flattened_frames_agree(product_lib, scene):
    """The same estimator over the two forms of the scene: hits are computed in object space in one and in world space in the
    other, so frames agree like the GPU and the oracle do, not bit for bit."""
    W, H, spp = 160, 120, 4
    g2, sc = product(product_lib, SCENES[scene], W, H, levels=2, spp=spp)
    g1, _ = product(product_lib, SCENES[scene], W, H, levels=1, spp=spp)
    assert "levels_in_use=1" in g1.get_settings()
    cam = sc.camera(W, H)
    g1.render_frame(cam, R.RESET), g2.render_frame(cam, R.RESET)
    a, b = g2.read_image(), g1.read_image()
    assert frac_bad(a, b) < (2e-2 if scene == "cornell" else 6e-2)
    assert abs(a[..., :3].mean() - b[..., :3].mean()) < 3e-3 * b[..., :3].mean() + 1e-4
    c2, c1 = g2.get_frame_counters().as_dict(), g1.get_frame_counters().as_dict()
    for k in ("n_ext", "n_shade", "n_nee"):
        assert abs(c2[k] - c1[k]) <= 2e-3 * c1[k] + 2


def test_two_level_settings_do_not_change_the_frame(product_lib):
    """re-ordering, batching and the (ignored) hit caches leave a two-level frame bit-identical"""
    W, H = 128, 96
    ref = None
    for extra in ({}, {"sort": "off"}, {"spp_batch": 1}, {"primary_cache": "off", "shadow_cache": "pixel"}):
        g, sc = product(product_lib, SCENES["soup"], W, H, levels=2, spp=3, **extra)
        g.render_frame(sc.camera(W, H), R.RESET)
        img = g.read_image().copy()
        if ref is None:
            ref = img
        assert np.array_equal(img, ref), extra


@pytest.mark.parametrize("scene", ["cornell", "soup"])
def test_both_two_level_kernels_render_the_same_frame(product_lib, scene):
    """Default: the while-while wavefront kernel with the instance step (packed nodes).  trace_variant 0: the plain
    one-ray-per-lane kernel around traverse_tl (fp32 nodes).  Same rays, same object-space triangle tests; packed boxes only
    contain the exact ones, so the frames differ at most where two triangles tie."""
    W, H, spp = 192, 128, 4
    a, sc = product(product_lib, SCENES[scene], W, H, levels=2, spp=spp)
    b, _ = product(product_lib, SCENES[scene], W, H, levels=2, spp=spp, trace_variant=0)
    cam = sc.camera(W, H)
    a.render_frame(cam, R.RESET), b.render_frame(cam, R.RESET)
    ia, ib = a.read_image(), b.read_image()
    assert (np.abs(ia - ib).max(axis=-1) > 0).mean() < 2e-3
    ca, cb = a.get_frame_counters().as_dict(), b.get_frame_counters().as_dict()
    for k in ("n_ext", "n_shade", "n_nee"):
        assert abs(ca[k] - cb[k]) <= 1e-3 * cb[k] + 2


def test_moving_an_instance_rebuilds_the_top_level_only(product_lib, oracle_lib):
    W, H = 160, 120
    g, sc = product(product_lib, SCENES["cornell"], W, H, levels=2)
    o = oracle(oracle_lib, SCENES["cornell"], W, H)
    # the six quads share one transform (one group, one top-level instance); the two blocks have their own
    info = dict(l.split("=", 1) for l in g.get_settings().splitlines() if "=" in l)
    assert int(info["instance_groups"]) == 3 and int(info["top_level_instances"]) == 3, info
    builds0 = g.get_geometry_stats().builds
    mesh, M = sc.instances[6]  # the short block
    M2 = S.translate(0.4, 0.3, -0.2) @ np.asarray(M)
    for ctx in (g, o):
        ctx.set_instance(6, mesh, M2)
        ctx.update()
    assert g.get_geometry_stats().builds == builds0  # no mesh tree was rebuilt
    origins, dirs = o.generate_primary(sc.camera(W, H), 1)
    hg, ho = g.trace_closest(origins, dirs), o.trace_closest(origins, dirs)
    assert ((ho["inst_id"] == 6) & (ho["prim_id"] >= 0)).sum() > 50
    _check_hits(g, o, origins, dirs, hg, ho)


def test_auto_switches_on_the_flatten_budget(product_lib):
    W, H = 96, 64
    g, sc = product(product_lib, SCENES["cornell"], W, H, levels="auto", flatten_budget=10)
    assert "levels_in_use=2" in g.get_settings()
    cam = sc.camera(W, H)
    g.render_frame(cam, R.RESET)
    a = g.read_image().copy()
    g.set_setting("flatten_budget", 1 << 26)  # now the scene fits: flattened again at the next update
    g.update()
    assert "levels_in_use=1" in g.get_settings()
    g.render_frame(cam, R.RESET)
    b = g.read_image()
    assert frac_bad(a, b) < 2e-2
    with pytest.raises(R.Rfwb200Error, match="levels"):
        g.set_setting("levels", "3")


def test_two_level_rejects_what_it_cannot_do(product_lib):
    g, sc = product(product_lib, SCENES["cornell"], 64, 64)
    g.set_setting("bvh", 8)
    g.set_setting("levels", 2)
    with pytest.raises(R.Rfwb200Error, match="levels=1"):
        g.update()


def test_config3_instanced_sponza_two_level_against_flattened_and_oracle(product_lib, oracle_lib):
    """The config-3 construction (the headline scene instanced on a lattice) at 4 copies: per-ray camera hits of the two-level
    scene against the oracle, and frame statistics against the flattened frame."""
    W, H, spp = 320, 180, 4
    fn = lambda: S.sponza_instanced(copies=4)
    g2, sc = product(product_lib, fn, W, H, levels=2, spp=spp)
    o = oracle(oracle_lib, fn, W, H)
    cam = sc.camera(W, H)
    origins, dirs = o.generate_primary(cam, 0)
    hg, ho = g2.trace_closest(origins, dirs), o.trace_closest(origins, dirs)
    same = (hg["inst_id"] == ho["inst_id"]) & (hg["prim_id"] == ho["prim_id"])
    assert same.mean() > 0.995, same.mean()
    hit = same & (ho["prim_id"] >= 0)
    assert (np.abs(hg["t"] - ho["t"])[hit] <= 2e-4 * np.maximum(1.0, ho["t"][hit])).all()
    g1, _ = product(product_lib, fn, W, H, levels=1, spp=spp)
    g1.render_frame(cam, R.RESET), g2.render_frame(cam, R.RESET)
    a, b = g2.read_image()[..., :3], g1.read_image()[..., :3]
    assert abs(a.mean() - b.mean()) < 0.01 * b.mean()
    blocks = lambda im: im[: H // 8 * 8, : W // 8 * 8].reshape(H // 8, 8, W // 8, 8, 3).mean(axis=(1, 3, 4))
    ba, bb = blocks(a), blocks(b)
    assert (np.abs(ba - bb) <= 0.15 * bb + 1e-3).mean() > 0.95
    i2, i1 = g2.get_bvh_info(), g1.get_bvh_info()
    assert i2["triangles"] < 0.4 * i1["triangles"]  # one copy of the records instead of four
    # the 393 meshes of the model are placed by the same four transforms: ONE tree below FOUR top-level instances (+ the light
    # quad), not 1,573 overlapping instance boxes
    info = dict(l.split("=", 1) for l in g2.get_settings().splitlines() if "=" in l)
    assert int(info["instance_groups"]) == 2 and int(info["top_level_instances"]) == 5, info


def test_groups_follow_the_placements(product_lib, oracle_lib):
    """Meshes placed by the same list of transforms share a tree; a mesh that is moved on its own leaves its group (both
    trees are rebuilt), and a mesh instanced twice gets two top-level instances over one tree.  Hits keep naming the caller's
    (instance, primitive) pairs."""
    W, H = 160, 120
    g, sc = product(product_lib, SCENES["cornell"], W, H, levels=2)
    o = oracle(oracle_lib, SCENES["cornell"], W, H)
    cam = sc.camera(W, H)
    mesh, M = sc.instances[2]  # the back wall, until now in the group of the six quads
    M2 = S.translate(0.0, 0.0, -0.5) @ np.asarray(M)
    extra = len(sc.instances)
    blk_mesh, blk_M = sc.instances[6]
    for ctx in (g, o):
        ctx.set_instance(2, mesh, M2)
        ctx.set_instance(extra, blk_mesh, S.translate(1.5, 0.0, -0.8) @ np.asarray(blk_M))  # the short block a second time
        ctx.update()
    info = dict(l.split("=", 1) for l in g.get_settings().splitlines() if "=" in l)
    assert int(info["instance_groups"]) == 4 and int(info["top_level_instances"]) == 5, info
    origins, dirs = o.generate_primary(cam, 0)
    hg, ho = g.trace_closest(origins, dirs), o.trace_closest(origins, dirs)
    assert ((ho["inst_id"] == extra) & (ho["prim_id"] >= 0)).sum() > 20 and ((ho["inst_id"] == 2) & (ho["prim_id"] >= 0)).sum() > 50
    _check_hits(g, o, origins, dirs, hg, ho)
    for ctx in (g, o):
        ctx.set_setting("spp", 1)
        ctx.set_probe_index(W // 2, H // 2)
        ctx.render_frame(cam, R.RESET)
    assert frac_bad(g.read_image(), o.read_image()) < 8e-3
    assert g.get_probe_results()[:2] == o.get_probe_results()[:2]


def test_two_level_scene_in_a_device_group(product_lib):
    """The in-process device group (here two ranks on device 0) over a two-level scene: every rank commits its own copy of the
    group trees; the sharded frame is bit-identical to the single-device two-level frame."""
    W, H = 200, 100
    one, sc = product(product_lib, SCENES["soup"], W, H, levels=2, spp=4)
    grp = R.RenderContext(product_lib, devices=[0, 0])
    grp.set_setting("levels", 2)
    S.upload(grp, SCENES["soup"](), W, H)
    grp.set_setting("spp", 4)
    assert "levels_in_use=2" in grp.get_settings()
    cam = sc.camera(W, H)
    for ctx in (one, grp):
        ctx.set_probe_index(W // 2, H // 2)
        ctx.render_frame(cam, R.RESET)
        ctx.render_frame(cam, R.CONVERGE)
    assert np.array_equal(one.read_image(), grp.read_image())
    assert one.get_probe_results() == grp.get_probe_results()
    assert one.get_frame_counters().as_dict() == grp.get_frame_counters().as_dict()
    grp.close()
