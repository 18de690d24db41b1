"""The scene ingest in front of the plugin boundary — Python producers in scenes.py / tools/bake_sponza.py —
pinned on the REFERENCE's own code: oracle/_ref/librfwref_ingest.so compiles texture::construct_mipmaps (texture.cpp:163-209), the
triangle LOD constant (geometry/assimp/object.cpp:728-731), the assimp material rule (material_list.cpp:54-78) and one light of
system::update_area_lights (system.cpp:1003-1025) from where they lie under /root/reference.  The committed vectors
(tests/golden/ref_ingest_vectors.npz, generator make_ref_ingest_golden.py) make the check run where the reference is absent;
where the library exists the vectors are re-derived from it first."""
import numpy as np
import pytest

import rfwb200 as R
import scenes as S
from ref_pin_ingest_common import GOLDEN, REF_INGEST_LIB, RefIngest, seeded_inputs


@pytest.fixture(scope="module")
def gold():
    g = dict(np.load(GOLDEN))
    if REF_INGEST_LIB.exists():  # the committed vectors are what the live reference code gives
        ref = RefIngest()
        texs, tris, mats = seeded_inputs()
        for i, t in enumerate(texs):
            assert np.array_equal(t, g[f"tex{i}"]) and np.array_equal(ref.mipmaps(t), g[f"mips{i}"])
        assert np.array_equal(np.array([ref.material(m) for m in mats], np.float32), g["mats_out"])
        assert np.array_equal(np.array([ref.lod(tris[i:i + 1], 256, 128) for i in range(len(tris))], np.float32), g["lod"], equal_nan=True)
    return g


def test_mip_chain_is_bit_identical(gold):
    for i in range(4):
        assert np.array_equal(S.build_mips(gold[f"tex{i}"]), gold[f"mips{i}"])


def test_triangle_lod_constant(gold):
    tris = gold["tris"]
    ref = gold["lod"]
    assert np.isfinite(ref).all() and (ref >= 0).all()  # max(0, sqrt(negative or log2(0))) is 0, never NaN
    assert np.allclose(tris["LOD"], ref, rtol=2e-6, atol=1e-6)
    assert (ref[:4] == 0).all() and (ref[4:] > 0).mean() > 0.5


def test_material_rule(gold):
    for vin, vout in zip(gold["mats_in"], gold["mats_out"]):
        m = S.material_rule(vin[0:3], vin[3:6], vin[6:9], vin[9], vin[10], vin[11], vin[12], vin[13])
        mine = np.concatenate([m["color"], m["absorption"], [m["metallic"], m["subsurface"], m["specular"], m["roughness"], m["eta"], m["transmission"]]]).astype(np.float32)
        assert np.allclose(mine, vout, rtol=1e-6, atol=1e-7), (vin, mine, vout)


def test_area_light_extraction_and_where_it_departs_from_the_reference(gold):
    """Identity instance: the producer equals system::update_area_lights field by field.  Transformed instances: the reference
    multiplies the row vector (vertex, 1) by the matrix (`vec4 * matrix`, math.h:878-940), i.e. by its TRANSPOSE — a translation is
    dropped and a rotation is inverted — and keeps the object-space area and the unnormalised normal; the producer places the
    light where the emissive triangle actually is (world-space vertices, unit normal, world-space area).  The test pins both:
    what the reference computes, and that the producer's lights sit on the instanced geometry."""
    tris, Ms, ref_lights = gold["tris"], gold["light_matrices"], gold["lights"]
    for k, M in enumerate(Ms):
        sc = S.Scene(name="lights")
        S.add_material(sc, (17.0, 12.0, 4.0))
        t = tris[8:16].copy()
        t["material"] = 0
        pos = np.stack([t["vertex0"], t["vertex1"], t["vertex2"]], 1)
        v4 = np.concatenate([pos.reshape(-1, 3), np.ones((24, 1), np.float32)], 1)
        sc.meshes = [S.SceneMesh(v4, t, None)]
        sc.instances = [(0, M)]
        S.extract_area_lights(sc)
        mine = sc.area_lights
        assert len(mine) == 8
        for i in range(8):
            r = ref_lights[8 * k + i]
            assert np.allclose(mine[i]["radiance"], r["radiance"]) and np.isclose(mine[i]["energy"], r["energy"], rtol=1e-6)
            assert mine[i]["tri_idx"] == r["tri_idx"] == i
            world = [(M[:3, :3] @ t[i][f].astype(np.float64) + M[:3, 3]) for f in ("vertex0", "vertex1", "vertex2")]
            for f, w in zip(("vertex0", "vertex1", "vertex2"), world):
                assert np.allclose(mine[i][f], w, atol=1e-4)  # on the instanced triangle
            if k == 0:  # identity: every field of the reference's light
                for f in ("position", "vertex0", "vertex1", "vertex2"):
                    assert np.allclose(mine[i][f], r[f], rtol=1e-6, atol=1e-6), f
                assert np.isclose(mine[i]["area"], r["area"], rtol=1e-5)
                n = r["normal"] / np.linalg.norm(r["normal"])
                assert np.allclose(mine[i]["normal"], n, atol=1e-5)
            else:  # the reference's row-vector product: transpose of the linear part, no translation
                expect = M[:3, :3].T @ t[i]["vertex0"].astype(np.float64)
                assert np.allclose(r["vertex0"], expect, atol=1e-4)
                assert np.isclose(r["area"], gold["light_tri_areas"][8 * k + i])  # object-space area (Triangle::updateArea)
