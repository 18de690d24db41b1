#!/usr/bin/env python3
"""Run seeded (instance, primitive, barycentrics) of the feature-soup scene through the REFERENCE's own E-mode material
lookup (oracle/_ref/librfwref_emat.so = Context::retrieve_material of EmbreeRT/src/Context.cpp:417-476 compiled from
/root/reference) and commit inputs + outputs as tests/golden/ref_emat_vectors.npz.  tests/test_ref_pin.py checks the
oracle's E-mode material step against them everywhere and against the live library where it exists."""
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(REPO / "rendering-fw_b200" / "python"))
sys.path.insert(0, str(REPO / "tests"))
import rfwb200 as R  # noqa: E402
from ref_pin_common import emat_scene, ref_emode_material  # noqa: E402

OUT = Path(__file__).resolve().parent / "ref_emat_vectors.npz"


def main():
    sc = emat_scene()
    rng = np.random.default_rng(4)
    rows_in, rows_out, kinds = [], [], []
    for _ in range(640):
        inst = int(rng.integers(0, len(sc.instances))) if _ % 8 else len(sc.instances) - 1  # every 8th: the FLOAT4-textured quad
        mi = sc.instances[inst][0]
        prim = int(rng.integers(0, len(sc.meshes[mi].triangles)))
        u = float(rng.random())
        v = float(rng.random() * (1 - u))
        color, N, iN = ref_emode_material(sc, inst, prim, u, v)
        rows_in.append((inst, prim, u, v)), rows_out.append(np.concatenate([color, N, iN]))
        m = sc.materials[sc.meshes[mi].triangles["material"][prim]]
        kind = 0
        if m["flags"] & R.MAT_HAS_DIFFUSE_MAP:
            kind = 1 if sc.textures[int(m["tex0"]["texaddr"])]["type"] == R.TEX_UINT else 2
        kinds.append(kind)
    np.savez_compressed(OUT, emat_in=np.array(rows_in, np.float64), emat_out=np.array(rows_out, np.float32), emat_kind=np.array(kinds, np.int32))
    print("wrote", OUT, "untextured / RGBA8 / float4:", [int((np.array(kinds) == k).sum()) for k in range(3)])


if __name__ == "__main__":
    main()
