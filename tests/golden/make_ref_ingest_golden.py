#!/usr/bin/env python3
"""Run seeded inputs through the REFERENCE's own ingest rules (oracle/_ref/librfwref_ingest.so: texture::construct_mipmaps, the
triangle LOD constant, the assimp material rule and one light of system::update_area_lights, compiled from /root/reference) and
commit inputs + reference outputs as tests/golden/ref_ingest_vectors.npz.  tests/test_ref_pin_ingest.py checks the Python
producers (scenes.build_mips / make_triangles / material_rule / extract_area_lights) against them everywhere and against the
live library where it exists.  Runs only in the build container."""
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(REPO / "rendering-fw_b200" / "python"))
sys.path.insert(0, str(REPO / "tests"))
import scenes as S  # noqa: E402
from ref_pin_ingest_common import GOLDEN, RefIngest, seeded_inputs  # noqa: E402


def main():
    ref = RefIngest()
    texs, tris, mats = seeded_inputs()
    out = {}
    for i, t in enumerate(texs):
        out[f"tex{i}"], out[f"mips{i}"] = t, ref.mipmaps(t)
    out["tris"] = tris
    out["lod"] = np.array([ref.lod(tris[i:i + 1], 256, 128) for i in range(len(tris))], np.float32)
    out["mats_in"] = mats
    out["mats_out"] = np.array([ref.material(m) for m in mats], np.float32)
    # area lights: identity, pure translation, rotation + non-uniform scale
    Ms = [np.eye(4), S.translate(3, 60, -2), S.translate(1, 2, 3) @ S.rotate_y(30.0) @ S.scale(2.0, 0.5, 1.5)]
    lights, areas = [], []
    for k, M in enumerate(Ms):
        for i in range(8):
            l, a = ref.area_light(tris[8 + i:9 + i], (17.0, 12.0, 4.0), M, i, k)
            lights.append(l), areas.append(a)
    out["light_matrices"] = np.array(Ms, np.float64)
    out["lights"] = np.array(lights)
    out["light_tri_areas"] = np.array(areas, np.float32)
    np.savez_compressed(GOLDEN, **out)
    print("wrote", GOLDEN, GOLDEN.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
