#!/usr/bin/env python3
"""Run seeded inputs through the REFERENCE's own shading / intersection headers (oracle/_ref/librfwref.so, built by
oracle/ref_build from the sources under /root/reference) and commit inputs + reference outputs as
tests/golden/ref_vectors.npz.  tests/test_ref_pin.py checks the oracle against these vectors everywhere, and against
the live library where it exists.  Runs only in the build container (the reference does not travel)."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(REPO / "rendering-fw_b200" / "python"))
sys.path.insert(0, str(REPO / "tests"))
import rfwb200 as R  # noqa: E402
import scenes as S  # noqa: E402
from ref_pin_common import RefLib, random_materials, soup_inputs  # noqa: E402

OUT = Path(__file__).resolve().parent / "ref_vectors.npz"


def main():
    ref = RefLib()
    rng = np.random.default_rng(2026)
    out = {}
    # 1-3: integer / packing
    seeds = rng.integers(0, 2 ** 32, size=64, dtype=np.uint64).astype(np.uint32)
    out["wang_in"], out["wang_out"] = seeds, np.array([ref.wang(int(s)) for s in seeds], np.uint32)
    out["rand_stream"] = ref.random_stream(0xC0FFEE, 16)
    n = rng.normal(size=(64, 3)).astype(np.float32)
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    out["pn_in"] = n
    out["pn_packed"] = np.array([ref.pack_normal(v) for v in n], np.uint32)
    out["pn_unpacked"] = np.array([ref.unpack_normal(int(p)) for p in out["pn_packed"]], np.float32)
    q = rng.integers(0, 300, size=(256, 4)).astype(np.int32)
    out["bn_in"], out["bn_out"] = q, np.array([ref.blue_noise(*map(int, r)) for r in q], np.float32)
    t_in = rng.normal(size=(64, 3)).astype(np.float32)
    t_in /= np.linalg.norm(t_in, axis=1, keepdims=True)
    out["ts_in"] = t_in
    out["ts_out"] = np.array([np.concatenate(ref.tangent_space(v)) for v in t_in], np.float32)
    # 4-5: Disney BSDF
    mats = random_materials(rng, 400)
    out.update({f"bsdf_{k}": v for k, v in mats.items()})
    ev = [ref.bsdf_eval(mats["color"][i], mats["params"][i], mats["N"][i], mats["wo"][i], mats["wi"][i]) for i in range(400)]
    out["bsdf_eval_out"] = np.array([np.concatenate([e[0], [e[1]]]) for e in ev], np.float32)
    sm = [ref.bsdf_sample(mats["color"][i], mats["absorption"][i], mats["params"][i], mats["N"][i], mats["wo"][i], float(mats["t"][i]),
                          int(mats["backfacing"][i]), float(mats["r3"][i]), float(mats["r4"][i])) for i in range(400)]
    out["bsdf_sample_out"] = np.array([np.concatenate([s[0], s[1], [s[2]]]) for s in sm], np.float32)
    # 6: triangle
    tri = rng.normal(size=(300, 3, 3)).astype(np.float32)
    org = (rng.normal(size=(300, 3)) * 2).astype(np.float32)
    tgt = (tri.mean(axis=1) + rng.normal(size=(300, 3)) * 0.4).astype(np.float32)
    d = tgt - org
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    out["tri_p"], out["tri_o"], out["tri_d"] = tri, org, d.astype(np.float32)
    out["tri_out"] = np.array([ref.intersect_triangle(org[i], d[i], 1e-5, 1e34, tri[i, 0], tri[i, 1], tri[i, 2], 1e-6) for i in range(300)], np.float32)
    # 7-9: soup scene (traversal, getShadingData, lights)
    out.update(soup_inputs(rng, ref=ref))
    np.savez_compressed(OUT, **out)
    print(OUT, OUT.stat().st_size)


if __name__ == "__main__":
    main()
