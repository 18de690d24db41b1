#!/usr/bin/env python3
"""Run seeded skinning and camera inputs through the REFERENCE's own SIMD math (oracle/_ref/librfwref_skin.so = rfw/math.h compiled
from /root/reference around the loop body of gltf/mesh.cpp:30-45) and commit inputs + reference outputs as
tests/golden/ref_skin_vectors.npz.  tests/test_ref_pin.py checks oracle/skinning.py against them everywhere and against
the live library where it exists.  Runs only in the build container (the reference does not travel)."""
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(REPO / "rendering-fw_b200" / "python"))
sys.path.insert(0, str(REPO / "tests"))
import scenes as S  # noqa: E402
from ref_pin_common import OracleScalar, emode_randoms, ref_camera_get_view, ref_generate_from_view, ref_set_pose  # noqa: E402

OUT = Path(__file__).resolve().parent / "ref_skin_vectors.npz"


def main():
    sc, sk = S.skinned_tube(seg=12, rings=10)
    rng = np.random.default_rng(77)
    pick = rng.choice(len(sk.base_vertices), size=96, replace=False)
    out = {"base_vertices": sk.base_vertices[pick], "base_normals": sk.base_normals[pick], "joints": sk.joints[pick], "weights": sk.weights[pick]}
    mats, vs, ns = [], [], []
    for k in (0, 9, 31, 58):
        J = sk.joint_matrices(k)
        if k == 58:  # a non-rigid pose as well: scale + shear on two joints
            J = J.copy()
            J[1] = J[1] @ np.diag([1.3, 0.8, 1.1, 1.0]).astype(np.float32)
            J[2, 0, 1] += 0.25
        v, n = ref_set_pose(J, out["base_vertices"], out["base_normals"], out["joints"], out["weights"])
        mats.append(J), vs.append(v), ns.append(n)
    out["joint_matrices"], out["ref_vertices"], out["ref_normals"] = np.array(mats, np.float32), np.array(vs), np.array(ns)
    # rfw::Camera::get_view (Camera.cpp:74-88) on seeded cameras: inputs (pos3, dir3, fov, focal, aperture, w, h) and the 14 outputs
    cams, views = [], []
    for _ in range(64):
        pos = rng.uniform(-300, 300, 3).astype(np.float32)
        d = rng.normal(size=3).astype(np.float32)
        d /= np.float32(np.linalg.norm(d))
        row = np.array([*pos, *d, rng.uniform(20, 100), rng.uniform(0.5, 20), rng.uniform(0, 0.1), rng.integers(16, 4000), rng.integers(16, 2200)], np.float32)
        cams.append(row)
        views.append(ref_camera_get_view(row[0:3], row[3:6], float(row[6]), float(row[7]), float(row[8]), int(row[9]), int(row[10])))
    out["camera_in"], out["camera_view"] = np.array(cams, np.float32), np.array(views, np.float32)
    # Ray::generateFromView (EmbreeRT/src/Ray.cpp:16-47) for pixels of a 96x64 frame under the E-mode random contract
    orc = OracleScalar()
    W, H, sample = 96, 64, 3
    view = views[0].copy()
    view[12] = 0.05  # a visible aperture so the lens term matters
    px = rng.choice(W * H, size=128, replace=False)
    rays = [ref_generate_from_view(view, W, H, int(p % W), int(p // W), emode_randoms(orc.wang, int(p), sample)) for p in px]
    out["ray_view"], out["ray_dims"], out["ray_pixels"], out["ray_out"] = view, np.array([W, H, sample], np.int32), px.astype(np.int32), np.array(rays, np.float32)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
