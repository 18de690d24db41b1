"""ctypes access to oracle/_ref/librfwref_ingest.so (four rules of the reference's scene ingest compiled from /root/reference,
oracle/ref_build/ref_ingest_shim.cpp) and the seeded inputs both the golden generator and the tests use."""
import ctypes as C
from pathlib import Path

import numpy as np

import rfwb200 as R

REPO = Path(__file__).resolve().parent.parent
REF_INGEST_LIB = REPO / "oracle" / "_ref" / "librfwref_ingest.so"
GOLDEN = Path(__file__).resolve().parent / "golden" / "ref_ingest_vectors.npz"


class RefIngest:
    def __init__(self):
        self.lib = C.CDLL(str(REF_INGEST_LIB))

    def mipmaps(self, level0: np.ndarray) -> np.ndarray:
        h, w = level0.shape
        n = sum((w >> l) * (h >> l) for l in range(5))
        out = np.zeros(n, np.uint32)
        f = self.lib.rfwref_construct_mipmaps
        f.restype = C.c_uint
        got = f(np.ascontiguousarray(level0, np.uint32).ctypes.data_as(C.c_void_p), C.c_uint(w), C.c_uint(h), out.ctypes.data_as(C.c_void_p), C.c_uint(n))
        assert got == n
        return out

    def lod(self, tri: np.ndarray, tw: int, th: int) -> float:
        f = self.lib.rfwref_triangle_lod
        f.restype = C.c_float
        t = np.ascontiguousarray(tri)
        return float(f(t.ctypes.data_as(C.c_void_p), C.c_uint(tw), C.c_uint(th)))

    def material(self, in14: np.ndarray) -> np.ndarray:
        out = np.zeros(12, np.float32)
        self.lib.rfwref_material_rule(np.ascontiguousarray(in14, np.float32).ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
        return out

    def area_light(self, tri: np.ndarray, color, matrix4: np.ndarray, index: int, inst: int):
        light = np.zeros(1, R.AREA_LIGHT_DTYPE)
        area = C.c_float()
        col = np.ascontiguousarray(color, np.float32)
        m = np.ascontiguousarray(np.asarray(matrix4, np.float32).T.reshape(-1))  # column-major, glm layout
        t = np.ascontiguousarray(tri)
        self.lib.rfwref_area_light(t.ctypes.data_as(C.c_void_p), col.ctypes.data_as(C.c_void_p), m.ctypes.data_as(C.c_void_p), C.c_int(index), C.c_int(inst),
                                   light.ctypes.data_as(C.c_void_p), C.byref(area))
        return light[0].copy(), float(area.value)


def seeded_inputs():
    """textures (odd and power-of-two sizes, alpha holes), triangles with uvs, assimp-style material values, emissive triangles"""
    import scenes as S

    rng = np.random.default_rng(2026)
    texs = []
    for h, w in ((64, 64), (48, 80), (16, 32), (33, 17)):
        rgba = rng.integers(0, 256, size=(h, w, 4), dtype=np.uint32)
        rgba[..., 3] = np.where(rng.random((h, w)) < 0.1, 0, 255)
        texs.append((rgba[..., 0] | (rgba[..., 1] << 8) | (rgba[..., 2] << 16) | (rgba[..., 3] << 24)).astype(np.uint32))
    pos = rng.uniform(-5, 5, size=(64, 3, 3)).astype(np.float32)
    uv = rng.uniform(-2, 3, size=(64, 3, 2)).astype(np.float32)
    uv[:4] = uv[:4, :1]  # degenerate uv triangles: Ta = 0 -> log2(0)
    tris = S.make_triangles(pos, None, uv, 0, tex_dims=(256, 128))
    mats = np.zeros((48, 14), np.float32)
    mats[:, 0:3] = np.where(rng.random((48, 1)) < 0.3, rng.uniform(0, 30, (48, 3)), 0)       # emissive
    mats[:, 3:6] = rng.uniform(-0.1, 1, (48, 3))                                               # diffuse (a negative component too)
    mats[:, 6:9] = rng.uniform(-0.2, 1, (48, 3))                                               # transparent
    mats[:, 9] = rng.choice([0.0, 1.0, 0.4, -0.5], 48)                                         # opacity
    mats[:, 10] = rng.choice([0.0, 10.0, 96.0, 1024.0, 5000.0], 48)                            # shininess
    mats[:, 11] = rng.choice([0.0, 0.3, 1.0, 2.0], 48)                                         # shininess strength
    mats[:, 12] = rng.choice([0.0, 1.0, 1.5], 48)                                              # eta
    mats[:, 13] = rng.choice([0.0, 0.7, -1.0], 48)                                             # reflectivity
    return texs, tris, mats
