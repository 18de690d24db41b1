"""CPU tests of the host-side logic around the boundary: camera, mips, area-light extraction, tile sharding,
baked-scene container, algorithmic-bytes formula."""
import math

import numpy as np
import pytest

import rfwb200 as R
import scenes as S


def test_camera_get_view_matches_reference_formula():
    # rfw::Camera::get_view (context/Camera.cpp:74-88): p1/p2/p3 span the focal plane, spread = FOV*pi/180/height
    cam = R.Camera((1, 2, 3), (0, 0, 1), fov=40.0, width=1920, height=1080)
    v = cam.get_view()
    assert abs(v.spread_angle - (40.0 * math.pi / 180) / 1080) < 1e-9
    p1, p2, p3 = np.array(v.p1[:]), np.array(v.p2[:]), np.array(v.p3[:])
    right, up = p2 - p1, p3 - p1
    s = math.tan(math.radians(20.0)) * 5.0
    assert np.allclose(np.linalg.norm(right), 2 * s * 1920 / 1080, rtol=1e-5)
    assert np.allclose(np.linalg.norm(up), 2 * s, rtol=1e-5)
    assert abs(np.dot(right, up)) < 1e-4
    centre = (p2 + p3) / 2
    assert np.allclose(centre, [1, 2, 8], atol=1e-4)  # position + focalDistance * forward
    assert up[1] < 0  # p3 is below p1: v grows downwards


def test_mip_chain_layout_and_box_filter():
    img = np.zeros((8, 8, 4), np.uint8)
    img[..., 0] = np.arange(64).reshape(8, 8)
    img[..., 3] = 255
    img[0, 0, 3] = 0
    sc = S.Scene()
    S.add_texture_rgba8(sc, img)
    t = sc.textures[0]
    assert t["data"].size == 64 + 16 + 4 + 1 + 0  # required_pixel_count (texture.cpp:211-225), 5 levels
    l1 = t["data"][64:80].reshape(4, 4)
    assert (l1[0, 0] & 255) == (0 + 1 + 8 + 9) >> 2
    assert (l1[0, 0] >> 24) == 0 and (l1[1, 1] >> 24) == 255  # alpha = min of the four texels


def test_area_light_extraction_follows_emissive_triangles():
    sc = S.cornell_box(unit_scale=True)
    S.extract_area_lights(sc)
    assert len(sc.area_lights) == 2
    light_mesh = sc.meshes[5]
    assert sorted(light_mesh.triangles["light_tri_idx"]) == [0, 1]
    assert all((m.triangles["light_tri_idx"] == -1).all() for i, m in enumerate(sc.meshes) if i != 5)
    L = sc.area_lights
    assert np.allclose(L["normal"], [[0, -1, 0]] * 2, atol=1e-6)
    assert np.allclose(L["area"].sum(), 1.30 * 1.05, rtol=1e-4)  # 130 x 105 quad at scale 0.01
    assert np.allclose(L["energy"], np.linalg.norm([17, 12, 4]), rtol=1e-3)
    assert np.allclose(L["position"], (L["vertex0"] + L["vertex1"] + L["vertex2"]) / 3, atol=1e-5)


def test_material_parameter_packing():
    sc = S.Scene()
    i = S.add_material(sc, (0.5, 0.25, 1.0), roughness=0.4, metallic=0.2, specular=0.6, subsurface=0.1, transmission=0.3, eta=1.5)
    m = sc.materials[i]
    p = m["parameters"]
    assert (p[0] & 255, (p[0] >> 8) & 255, (p[0] >> 16) & 255, p[0] >> 24) == (51, 25, 153, 102)  # uint(x*255), material_list.cpp:337
    assert (p[2] >> 16) & 255 == 76 and p[2] >> 24 == int(0.75 * 255)
    assert np.allclose(m["diffuse"].astype(np.float32), [0.5, 0.25, 1.0])


@pytest.mark.parametrize("w,h,world,tw,th", [(1920, 1080, 1, 32, 8), (1920, 1080, 8, 32, 8), (200, 100, 3, 32, 8), (64, 64, 2, 64, 16), (33, 7, 4, 8, 4)])
def test_tile_shards_partition_the_image(w, h, world, tw, th):
    seen = np.zeros(w * h, np.int32)
    for r in range(world):
        m = R.shard_pixel_map(w, h, r, world, tw, th)
        assert len(m) <= R.shard_stride(w, h, world, tw, th)
        ok = m[m >= 0]
        seen[ok] += 1
        # a warp's 32 consecutive work items form an 8x4 pixel block
        blk = m[:32]
        if (blk >= 0).all():
            xs, ys = blk % w, blk // w
            assert xs.max() - xs.min() == 7 and ys.max() - ys.min() == 3
    assert (seen == 1).all()


def test_assemble_shards_host_roundtrip():
    w, h, world = 200, 100, 3
    img = np.random.default_rng(0).random((h * w, 4)).astype(np.float32)
    shards = []
    for r in range(world):
        m = R.shard_pixel_map(w, h, r, world)
        s = np.zeros((R.shard_stride(w, h, world), 4), np.float32)
        s[: len(m)][m >= 0] = img[m[m >= 0]]
        shards.append(s)
    assert np.array_equal(R.assemble_shards_host(shards, w, h).reshape(-1, 4), img)


def test_baked_scene_roundtrip(tmp_path):
    sc = S.feature_soup(200)
    S.extract_area_lights(sc)
    S.save_baked(sc, tmp_path / "s.rfwscene")
    b = S.load_baked(tmp_path / "s.rfwscene")
    assert len(b.meshes) == len(sc.meshes) and len(b.instances) == len(sc.instances)
    assert all(np.array_equal(x.triangles, y.triangles) and np.array_equal(x.vertices, y.vertices) for x, y in zip(b.meshes, sc.meshes))
    assert np.array_equal(b.materials, sc.materials) and np.array_equal(b.tex_ids, sc.tex_ids)
    assert all(np.array_equal(x["data"], y["data"]) for x, y in zip(b.textures, sc.textures))
    assert np.array_equal(b.point_lights, sc.point_lights) and tuple(b.camera_pos) == pytest.approx(sc.camera_pos)


def test_algorithmic_bytes_formula():
    fc = R.FrameCounters(n_gen=10, n_ext=25, n_shade=25, n_ext_out=15, n_nee=12, n_acc=9, pixels=10, samples=1)
    # B = N_gen*32 + N_ext*48 + N_shade*224 + N_extOut*48 + N_nee*96 + N_acc*32 + P*32   
    assert fc.algorithmic_bytes() == 10 * 32 + 25 * 48 + 25 * 224 + 15 * 48 + 12 * 96 + 9 * 32 + 10 * 32


def test_atrium_standin_is_well_formed():
    sc = S.atrium(target_tris=20000)
    assert 10000 < sc.triangle_count() < 60000
    S.extract_area_lights(sc)
    assert len(sc.area_lights) == 2
    for m in sc.meshes:
        assert np.isfinite(m.vertices).all() and np.isfinite(m.triangles["LOD"]).all()


def test_closed_form_random_barycentrics_equals_the_reference_loop(oracle_lib):
    """k_shade's RandomBarycentrics replaces the sixteen-trip subdivision loop of lights.h:119-157 by the weights of the
    final corners in 16.16 fixed point (csrc/kernels.cu).  This numpy restatement of those integer sums must give the
    oracle's loop (pinned on the reference's own lights.h, test_ref_pin) bit for bit — every float of the loop is exact."""
    import ctypes as C

    f = oracle_lib.fn("random_barycentrics", None, [C.c_float, C.c_void_p])
    rng = np.random.default_rng(7)
    r0s = np.concatenate([rng.random(4000, dtype=np.float32), np.float32([0.0, 0.25, 0.5, 0.75, 1.0 - 2 ** -24, 1e-9, 0.3333333])])
    for r0 in r0s:
        out = np.zeros(3, np.float32)
        f(float(r0), out.ctypes.data)
        uf = min(int(np.float32(r0) * np.float32(4294967295.0)), 0xFFFFFFFF)  # __float2uint_rz saturates
        H, wA, wB = 3 << 15, 1 << 16, 1 << 16
        for _ in range(16):
            d, hA, hB = uf & 3, wA >> 1, wB >> 1
            wA = H - hA if d == 0 else (hA + H if d == 1 else hA)
            wB = H - hB if d == 0 else (hB + H if d == 2 else hB)
            uf >>= 2
        rx = np.float32(np.float32(wA) * np.float32(1.0 / 65536.0)) * np.float32(0.3333333)
        ry = np.float32(np.float32(wB) * np.float32(1.0 / 65536.0)) * np.float32(0.3333333)
        mine = np.float32([rx, ry, np.float32(np.float32(1.0) - rx) - ry])
        assert np.array_equal(mine, out), (r0, mine, out)
