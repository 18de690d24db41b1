"""CPU tests of the oracle itself (it is the checker, so it is checked first): known-answer tests for the
integer / bit-exact pieces against independent Python restatements, the acceleration structure against brute
force, estimator sanity, and the committed golden fixtures (tests/golden/make_golden.py)."""
import ctypes as C
from pathlib import Path

import numpy as np
import pytest

import rfwb200 as R
import scenes as S

GOLD = Path(__file__).resolve().parent / "golden"


def py_wang(s):
    s &= 0xFFFFFFFF
    s = ((s ^ 61) ^ (s >> 16)) & 0xFFFFFFFF
    s = (s * 9) & 0xFFFFFFFF
    s = s ^ (s >> 4)
    s = (s * 0x27D4EB2D) & 0xFFFFFFFF
    return s ^ (s >> 15)


def test_wang_hash_and_xorshift(oracle_lib):
    wang = oracle_lib.fn("wang_hash", C.c_uint32, [C.c_uint32])
    for s in (0, 1, 61, 12345, 0xDEADBEEF, 0xFFFFFFFF):
        assert wang(s) == py_wang(s)
    st = C.c_uint32(0x9E3779B9)
    rint = oracle_lib.fn("random_int", C.c_uint32, [C.c_void_p])
    x = 0x9E3779B9
    for _ in range(6):
        x ^= (x << 13) & 0xFFFFFFFF
        x ^= x >> 17
        x ^= (x << 5) & 0xFFFFFFFF
        assert rint(C.byref(st)) == x


def test_xor128_known_answers(oracle_lib):
    # Marsaglia's xorshift128 with the default state of rfw::utils::xor128 (utils/xor128.h:30-33)
    xor128 = oracle_lib.fn("xor128", C.c_uint32, [C.c_uint32, C.c_uint32])
    assert [xor128(123456789, n) for n in (1, 2, 3)] == [3701687786, 458299110, 2500872618]


def test_half_conversion_matches_numpy(oracle_lib):
    h2f = oracle_lib.fn("half_to_float", C.c_float, [C.c_uint16])
    bits = np.concatenate([np.arange(0, 0x7C00, 97), np.array([0, 1, 0x3FF, 0x400, 0x3C00, 0x7BFF, 0x8000, 0xBC00, 0x7C00])]).astype(np.uint16)
    ref = bits.view(np.float16).astype(np.float32)
    got = np.array([h2f(int(b)) for b in bits], np.float32)
    assert np.array_equal(got, ref)


def test_blue_noise_sampler_against_table(oracle_lib):
    table = np.fromfile(R.BLUENOISE_BIN, dtype=np.uint8)
    assert table.size == 327680
    ctx = R.RenderContext(oracle_lib)
    bn = oracle_lib.fn("blue_noise", C.c_float, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int])
    rng = np.random.default_rng(0)
    for x, y, s, d in rng.integers(0, 400, size=(200, 4)):
        xx, yy, ss, dd = x & 127, y & 127, s & 255, d & 255
        ranked = ss ^ int(table[dd + (xx + yy * 128) * 8 + 65536 * 3])
        v = int(table[dd + ranked * 256]) ^ int(table[(dd & 7) + (xx + yy * 128) * 8 + 65536])
        assert bn(ctx._h, int(x), int(y), int(s), int(d)) == np.float32((0.5 + v) * (1.0 / 256.0))


def test_pack_unpack_normal_roundtrip(oracle_lib):
    pack = oracle_lib.fn("pack_normal", C.c_uint32, [C.c_void_p])
    unpack = oracle_lib.fn("unpack_normal", None, [C.c_uint32, C.c_void_p])
    rng = np.random.default_rng(2)
    n = rng.normal(size=(200, 3)).astype(np.float32)
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    out = np.zeros(3, np.float32)
    for v in n:
        if v[2] < -0.8:  # 16-bit spheremap loses precision towards -z
            continue
        unpack(pack(v.ctypes.data), out.ctypes.data)
        assert np.abs(out - v).max() < 3e-4


def test_random_barycentrics_is_a_valid_point(oracle_lib):
    f = oracle_lib.fn("random_barycentrics", None, [C.c_float, C.c_void_p])
    out = np.zeros(3, np.float32)
    for r in np.linspace(0, 0.999, 64):
        f(float(r), out.ctypes.data)
        assert abs(out.sum() - 1) < 1e-5 and (out >= -1e-6).all()


def brute_force_closest(scene, origins, dirs, t_min=1e-5):
    """O(rays x triangles) double-precision Moller-Trumbore over the flattened scene."""
    tris = []
    for ii, (mi, M) in enumerate(scene.instances):
        m = scene.meshes[mi]
        v = m.vertices[:, :3].astype(np.float64)
        idx = m.indices if m.indices is not None else np.arange(len(m.triangles) * 3).reshape(-1, 3)
        w = v[idx] @ np.asarray(M)[:3, :3].T + np.asarray(M)[:3, 3]
        for pi in range(len(idx)):
            tris.append((ii, pi, w[pi]))
    P = np.array([t[2] for t in tris])
    ids = np.array([(t[0], t[1]) for t in tris])
    best_t = np.full(len(origins), 1e34)
    best = np.full((len(origins), 2), -1)
    e1, e2 = P[:, 1] - P[:, 0], P[:, 2] - P[:, 0]
    for r in range(len(origins)):
        o, d = origins[r, :3].astype(np.float64), dirs[r, :3].astype(np.float64)
        h = np.cross(d, e2)
        a = np.einsum("ij,ij->i", e1, h)
        ok = np.abs(a) > 1e-12
        f = np.where(ok, 1.0 / np.where(ok, a, 1), 0)
        s = o - P[:, 0]
        u = f * np.einsum("ij,ij->i", s, h)
        q = np.cross(s, e1)
        v = f * (q @ d)
        t = f * np.einsum("ij,ij->i", e2, q)
        hit = ok & (u >= 0) & (u <= 1) & (v >= 0) & (u + v <= 1) & (t > t_min)
        if hit.any():
            k = np.argmin(np.where(hit, t, np.inf))
            best_t[r], best[r] = t[k], ids[k]
    return best_t, best


@pytest.mark.parametrize("scene_fn", [lambda: S.cornell_box(unit_scale=True), lambda: S.feature_soup(600)])
def test_two_level_mbvh_against_brute_force(oracle_lib, scene_fn):
    sc = scene_fn()
    ctx = R.RenderContext(oracle_lib)
    S.upload(ctx, sc, 32, 24)
    o, d = ctx.generate_primary(sc.camera(32, 24), 0)
    hits = ctx.trace_closest(o, d)
    bt, bid = brute_force_closest(sc, o, d)
    hit = bid[:, 1] >= 0
    assert ((hits["prim_id"] >= 0) == hit).mean() > 0.995
    both = hit & (hits["prim_id"] >= 0)
    assert np.abs(hits["t"][both] - bt[both]).max() < 1e-3
    same = (hits["inst_id"] == bid[:, 0]) & (hits["prim_id"] == bid[:, 1])
    assert same[both].mean() > 0.98  # ties on shared edges may pick the neighbour


def test_emode_and_pt_agree_on_visibility(oracle_lib):
    sc = S.cornell_box(unit_scale=True)
    ctx = R.RenderContext(oracle_lib)
    S.upload(ctx, sc, 48, 48)
    cam = sc.camera(48, 48)
    ctx.set_setting("mode", "embree")
    ctx.render_frame(cam, R.RESET)
    e = ctx.read_image().copy()
    ctx.set_setting("mode", "pt")
    ctx.set_setting("spp", 16)
    ctx.render_frame(cam, R.RESET)
    p = ctx.read_image().copy()
    assert (e[..., 3] == 1).mean() > 0.8  # the open front of the box fills most of the view
    assert np.isfinite(p).all() and p[..., :3].mean() > 0.05
    # the light is seen directly in both models
    light = e[..., 0] > 5
    assert light.sum() > 4 and (p[light][:, 0] > 5).mean() > 0.6  # edge pixels differ: the two modes jitter differently


def test_pt_estimator_is_deterministic_and_sample_dependent(oracle_lib):
    sc = S.cornell_box(unit_scale=True)
    imgs = []
    for spp in (1, 1, 2):
        ctx = R.RenderContext(oracle_lib)
        S.upload(ctx, sc, 40, 30)
        ctx.set_setting("spp", spp)
        ctx.render_frame(sc.camera(40, 30), R.RESET)
        imgs.append(ctx.read_image().copy())
    assert np.array_equal(imgs[0], imgs[1])
    assert not np.array_equal(imgs[0], imgs[2])


def test_thread_count_does_not_change_the_image(oracle_lib):
    sc = S.feature_soup(400)
    imgs = []
    for threads in (1, 5):
        ctx = R.RenderContext(oracle_lib)
        S.upload(ctx, sc, 40, 30)
        ctx.set_setting("threads", threads)
        ctx.set_setting("spp", 2)
        ctx.render_frame(sc.camera(40, 30), R.RESET)
        imgs.append(ctx.read_image().copy())
    ctx.set_setting("threads", 0 if False else 8)
    assert np.array_equal(imgs[0], imgs[1])


def test_golden_scalar_kats(oracle_lib):
    g = np.load(GOLD / "scalar_kats.npz")
    wang = oracle_lib.fn("wang_hash", C.c_uint32, [C.c_uint32])
    assert [wang(int(s)) for s in g["wang_in"]] == list(g["wang_out"])
    assert list(g["xor128_default_first8"][:3]) == [3701687786, 458299110, 2500872618]
    pack = oracle_lib.fn("pack_normal", C.c_uint32, [C.c_void_p])
    n = g["normals"]
    assert [pack(n[i].ctypes.data) for i in range(len(n))] == list(g["packed"])


@pytest.mark.parametrize("scene", ["cornell", "soup"])
def test_golden_frames(oracle_lib, scene):
    g = np.load(GOLD / "oracle_frames_64x48.npz")
    sc = S.cornell_box(unit_scale=True) if scene == "cornell" else S.feature_soup()
    ctx = R.RenderContext(oracle_lib)
    W, H = 64, 48
    S.upload(ctx, sc, W, H)
    cam = sc.camera(W, H)
    o, d = ctx.generate_primary(cam, 0)
    assert np.array_equal(o, g[f"{scene}_origins"]) and np.array_equal(d, g[f"{scene}_dirs"])
    hits = ctx.trace_closest(o, d)
    gh = g[f"{scene}_hits"]
    assert np.array_equal(hits["prim_id"], gh["prim_id"]) and np.array_equal(hits["inst_id"], gh["inst_id"])
    assert np.allclose(hits["t"], gh["t"], rtol=1e-6)
    for mode, depth in (("embree", 2), ("pt", 0), ("pt", 1), ("pt", 2)):
        ctx.set_setting("mode", mode)
        ctx.set_setting("max_path_length", depth)
        ctx.set_setting("spp", 1)
        ctx.render_frame(cam, R.RESET)
        img = ctx.read_image()
        ref = g[f"{scene}_{mode}_d{depth}"]
        # same compiler, same flags => identical; allow libm ulp drift across glibc versions
        assert np.allclose(img, ref, rtol=2e-4, atol=2e-5), (mode, depth, np.abs(img - ref).max())
        counters = np.array(list(ctx.get_frame_counters().as_dict().values()), np.uint64)
        assert np.abs(counters.astype(np.int64) - g[f"{scene}_{mode}_d{depth}_counters"].astype(np.int64)).max() <= 2


# ---- skinning restatement (oracle/skinning.py; gltf/mesh.cpp:18-48, 428-449) ---------------------------------
def test_skinning_restatement_known_answers():
    from oracle import skinning as K

    sc, sk = S.skinned_tube()
    m = sc.meshes[sk.mesh_index]
    # identity joints: the bind pose comes back
    I = np.broadcast_to(np.eye(4, dtype=np.float32), (sk.n_joints, 4, 4))
    v, n = K.set_pose(sk.base_vertices, sk.base_normals, sk.joints, sk.weights, I)
    assert np.allclose(v, sk.base_vertices, atol=1e-6) and np.allclose(n, sk.base_normals, atol=1e-6)
    t = K.update_triangles(m.triangles, v, n, m.indices)
    assert np.allclose(t["vertex1"], m.triangles["vertex1"], atol=1e-6)
    # one rotation on every joint: vertices move rigidly, normals rotate, lengths stay 1
    Rm = S.rotate_y(33.0).astype(np.float32)
    v, n = K.set_pose(sk.base_vertices, sk.base_normals, sk.joints, sk.weights, np.broadcast_to(Rm, (sk.n_joints, 4, 4)))
    assert np.allclose(v, sk.base_vertices @ Rm.T, atol=2e-6)
    assert np.allclose(n, sk.base_normals @ Rm[:3, :3].T, atol=2e-6)
    assert np.allclose(np.linalg.norm(n, axis=1), 1.0, atol=1e-6)
    # with a translation the reference's vec4 length (math.h:797-806) includes w = n . t_inv: direction right, length < 1
    Tm = (S.translate(0.3, -0.1, 0.2) @ S.rotate_y(33.0)).astype(np.float32)
    v, n = K.set_pose(sk.base_vertices, sk.base_normals, sk.joints, sk.weights, np.broadcast_to(Tm, (sk.n_joints, 4, 4)))
    assert np.allclose(v, sk.base_vertices @ Tm.T, atol=2e-6)
    rot = sk.base_normals @ Tm[:3, :3].T
    ln = np.linalg.norm(n, axis=1, keepdims=True)
    assert np.allclose(n / ln, rot, atol=2e-6) and (ln <= 1.0 + 1e-6).all() and ln.min() < 0.999
    # non-uniform scale: normals follow the inverse transpose
    Sm = np.diag([2.0, 0.5, 1.0, 1.0]).astype(np.float32)
    v, n = K.set_pose(sk.base_vertices, sk.base_normals, sk.joints, sk.weights, np.broadcast_to(Sm, (sk.n_joints, 4, 4)))
    expect = sk.base_normals @ np.linalg.inv(Sm[:3, :3])
    expect /= np.linalg.norm(expect, axis=1, keepdims=True)
    assert np.allclose(n, expect, atol=2e-6)
    # the animated pose keeps the surface closed: geometric normals agree with the skinned vertex normals
    v, n = K.set_pose(sk.base_vertices, sk.base_normals, sk.joints, sk.weights, sk.joint_matrices(40))
    t = K.update_triangles(m.triangles, v, n, m.indices)
    gn = np.stack([t["Nx"], t["Ny"], t["Nz"]], 1)
    assert (np.einsum("ij,ij->i", gn, t["vN0"]) > 0.5).mean() > 0.99


def test_tone_map_restatement_against_the_shader_formulas():
    """oracle/tonemap.py: the float32 restatement of assets/shaders/tone-map.frag stays within half a code value of the float64
    evaluation of the same formulas; anchors computed by hand from the shader."""
    import sys

    sys.path.insert(0, str(R.REPO_DIR))
    from oracle.tonemap import tone_map, tone_map_spec

    rng = np.random.default_rng(5)
    x = np.concatenate([rng.uniform(0, 6, (4000, 4)), rng.uniform(0, 0.1, (1000, 4)), [[0, 0, 0, 0], [1e9, 1e9, 1e9, 2], [17, 12, 4, 1]]]).astype(np.float32)
    for contrast, brightness in ((1.0, 0.05), (0.0, 0.0), (1.7, -0.2)):
        b, s = tone_map(x, contrast, brightness), tone_map_spec(x, contrast, brightness)
        assert b.dtype == np.uint8 and np.abs(b.astype(np.float64) - s).max() <= 0.5 + 1e-3
    # black with contrast 1, brightness 0: rgb = 0 -> RRTAndODTFit(0) < 0 -> clamped to 0; alpha passes through
    assert tone_map(np.array([[0, 0, 0, 0.5]], np.float32), 1.0, 0.0).tolist() == [[0, 0, 0, 128]]
    # mid grey 0.18 with the camera defaults (contrast 1, brightness 0.05): v = 0.23 per channel -> ACES ~ 0.3217 -> code 82
    v = 0.23 * (0.59719 + 0.35458 + 0.04823)
    fit = (v * (v + 0.0245786) - 0.000090537) / (v * (0.983729 * v + 0.4329510) + 0.238081)
    code = round(255 * fit * (1.60475 - 0.53108 - 0.07367))
    assert abs(int(tone_map(np.array([[0.18, 0.18, 0.18, 1]], np.float32), 1.0, 0.05)[0, 0]) - code) <= 1
    # monotone in luminance, saturates at white
    ramp = np.linspace(0, 50, 256, dtype=np.float32)[:, None].repeat(4, 1)
    t = tone_map(ramp, 1.0, 0.05)[:, 0].astype(int)
    assert (np.diff(t) >= 0).all() and t[-1] == 255
