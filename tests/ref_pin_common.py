"""Shared helpers of tests/golden/make_ref_golden.py and tests/test_ref_pin.py: ctypes wrappers around
oracle/_ref/librfwref.so (the reference's headers) and around the matching oracle hooks, plus input generators."""
import ctypes as C
from pathlib import Path

import numpy as np

import rfwb200 as R
import scenes as S

REF_LIB = R.REPO_DIR / "oracle" / "_ref" / "librfwref.so"
REF_SKIN_LIB = R.REPO_DIR / "oracle" / "_ref" / "librfwref_skin.so"
F, U, I, P = C.c_float, C.c_uint32, C.c_int, C.c_void_p


REF_CAMERA_LIB = R.REPO_DIR / "oracle" / "_ref" / "librfwref_camera.so"


def ref_camera_get_view(position, direction, fov, focal_distance, aperture, width, height):
    """rfw::Camera::get_view of the reference's own Camera.cpp (oracle/ref_build/ref_camera_shim.cpp) -> 14 floats"""
    lib = C.CDLL(str(REF_CAMERA_LIB))
    f = lib.rfwref_camera_get_view
    f.restype, f.argtypes = None, [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_void_p]
    p, d = np.ascontiguousarray(position, np.float32), np.ascontiguousarray(direction, np.float32)
    out = np.zeros(14, np.float32)
    f(p.ctypes.data, d.ctypes.data, fov, focal_distance, aperture, width, height, out.ctypes.data)
    return out


REF_RAY_LIB = R.REPO_DIR / "oracle" / "_ref" / "librfwref_ray.so"


def ref_generate_from_view(view14, width, height, x, y, r):
    """Ray::generateFromView of the reference (EmbreeRT/src/Ray.cpp:16-47, oracle/ref_build/ref_ray_shim.cpp) -> origin, direction"""
    lib = C.CDLL(str(REF_RAY_LIB))
    f = lib.rfwref_generate_from_view
    f.restype, f.argtypes = None, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int] + [C.c_float] * 4 + [C.c_void_p]
    v = np.ascontiguousarray(view14, np.float32)
    out = np.zeros(6, np.float32)
    f(v.ctypes.data, width, height, x, y, float(r[0]), float(r[1]), float(r[2]), float(r[3]), out.ctypes.data)
    return out


def emode_randoms(wang, pixel, sample):
    """the E-mode determinism contract: first four xor128 outputs with x seeded per (pixel, sample)"""
    m = 0xFFFFFFFF
    x, y, z, w = 123456789 ^ wang((pixel * 16789 + sample * 1791) & m), 362436069, 521288629, 88675123
    out = []
    for _ in range(4):
        t = (x ^ (x << 11)) & m
        x, y, z = y, z, w
        w = (w ^ (w >> 19) ^ (t ^ (t >> 8))) & m
        out.append(np.float32(w) * np.float32(2.3283064365387e-10))
    return np.array(out, np.float32)


def ref_set_pose(joint_matrices, base_vertices, base_normals, joints, weights):
    """the reference's own SIMD math around the loop body of gltf/mesh.cpp:30-45 (oracle/ref_build/ref_skin_shim.cpp).
    joint_matrices: (nj, 4, 4) in the mathematical (row, col) convention -> (vertices (nv,4), normals (nv,3))"""
    lib = C.CDLL(str(REF_SKIN_LIB))
    f = lib.rfwref_set_pose
    f.restype, f.argtypes = None, [C.c_void_p] * 5 + [C.c_int] + [C.c_void_p] * 2
    J = np.ascontiguousarray(np.transpose(np.asarray(joint_matrices, np.float32), (0, 2, 1)))  # column-major like glm
    bv = np.ascontiguousarray(base_vertices, np.float32).reshape(-1, 4)
    nv = len(bv)
    bn = np.zeros((nv, 4), np.float32)
    bn[:, :3] = np.asarray(base_normals, np.float32).reshape(nv, -1)[:, :3]
    j = np.ascontiguousarray(joints, np.uint32).reshape(nv, 4)
    w = np.ascontiguousarray(weights, np.float32).reshape(nv, 4)
    ov, on = np.zeros((nv, 4), np.float32), np.zeros((nv, 3), np.float32)
    f(J.ctypes.data, bv.ctypes.data, bn.ctypes.data, j.ctypes.data, w.ctypes.data, nv, ov.ctypes.data, on.ctypes.data)
    return ov, on


def fp(a):
    return np.ascontiguousarray(a, np.float32).ctypes.data


def blue_noise_uint_table():
    t = np.fromfile(R.BLUENOISE_BIN, dtype=np.uint8)
    buf = np.zeros(65536 * 5, np.uint32)  # createBlueNoiseBuffer, blue_noise.h:8204-8219
    buf[:65536] = t[:65536]
    buf[65536:65536 + 131072] = t[65536:65536 + 131072]
    buf[3 * 65536:3 * 65536 + 131072] = t[65536 + 131072:]
    return buf


class RefLib:
    """the reference's own code (oracle/_ref)"""

    prefix = "rfwref_"

    def __init__(self, path=REF_LIB):
        self.lib = C.CDLL(str(path))
        self.bn = blue_noise_uint_table()

    def f(self, name, res, args):
        fn = getattr(self.lib, self.prefix + name)
        fn.restype, fn.argtypes = res, args
        return fn

    def wang(self, s):
        return self.f("wang_hash", U, [U])(s)

    def random_stream(self, seed, n):
        st = U(seed)
        fn = self.f("random_int", U, [P])
        return np.array([fn(C.byref(st)) for _ in range(n)], np.uint32)

    def pack_normal(self, v):
        return self.f("pack_normal", U, [P])(fp(v))

    def unpack_normal(self, p):
        o = np.zeros(3, np.float32)
        self.f("unpack_normal", None, [U, P])(p, o.ctypes.data)
        return o

    def blue_noise(self, x, y, s, d):
        return self.f("blue_noise", F, [P, I, I, I, I])(self.bn.ctypes.data, x, y, s, d)

    def tangent_space(self, n):
        T, B = np.zeros(3, np.float32), np.zeros(3, np.float32)
        self.f("tangent_space", None, [P, P, P])(fp(n), T.ctypes.data, B.ctypes.data)
        return T, B

    def bsdf_eval(self, color, params, N, wo, wi):
        o, pdf = np.zeros(3, np.float32), F()
        prm = np.ascontiguousarray(params, np.uint32)
        self.f("bsdf_eval", None, [P] * 6 + [P])(fp(color), prm.ctypes.data, fp(N), fp(wo), fp(wi), o.ctypes.data, C.addressof(pdf))
        return o, pdf.value

    def bsdf_sample(self, color, absorption, params, N, wo, t, backfacing, r3, r4):
        wi, o, pdf = np.zeros(3, np.float32), np.zeros(3, np.float32), F()
        prm = np.ascontiguousarray(params, np.uint32)
        name = "bsdf_sample" if self.prefix == "rfwref_" else "bsdf_sample_r"
        self.f(name, None, [P, P, P, P, P, F, I, F, F, P, P, P])(fp(color), fp(absorption), prm.ctypes.data, fp(N), fp(wo), t, backfacing,
                                                                  r3, r4, wi.ctypes.data, o.ctypes.data, C.addressof(pdf))
        return wi, o, pdf.value

    def intersect_triangle(self, o, d, tmin, tmax, p0, p1, p2, eps):
        t, b = F(), np.zeros(2, np.float32)
        hit = self.f("intersect_triangle", I, [P, P, F, F, P, P, P, F, P, P])(fp(o), fp(d), tmin, tmax, fp(p0), fp(p1), fp(p2), eps,
                                                                               C.addressof(t), b.ctypes.data)
        return hit, t.value, b[0], b[1]

    def random_barycentrics(self, r0):
        o = np.zeros(3, np.float32)
        self.f("random_barycentrics", None, [F, P])(r0, o.ctypes.data)
        return o


class OracleScalar(RefLib):
    """the same call shapes on the oracle (oracle/librfworacle.so)"""

    prefix = "rfworacle_"

    def __init__(self):
        self.L = R.load_oracle()
        self.lib = self.L.lib
        self.ctx = R.RenderContext(self.L)

    def blue_noise(self, x, y, s, d):
        return self.f("blue_noise", F, [P, I, I, I, I])(self.ctx._h, x, y, s, d)

    def bsdf_eval(self, color, params, N, wo, wi):
        o, pdf = np.zeros(3, np.float32), F()
        prm = np.ascontiguousarray(params, np.uint32)
        self.f("bsdf_eval", None, [P] * 7)(fp(color), prm.ctypes.data, fp(N), fp(wo), fp(wi), o.ctypes.data, C.addressof(pdf))
        return o, pdf.value


def random_materials(rng, n):
    def unit(v):
        return (v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float32)

    N = unit(rng.normal(size=(n, 3)))
    wo = unit(N + rng.normal(size=(n, 3)) * 0.8)
    flipw = np.einsum("ij,ij->i", wo, N) < 0.05
    wo[flipw] = unit(N[flipw] * 1.5 + wo[flipw])
    wi = unit(rng.normal(size=(n, 3)))
    params = rng.integers(0, 256, size=(n, 16)).astype(np.uint32)
    params[: n // 3, 10] = 0  # transmission = 0 for a third
    params[n // 3: n // 2, 1] = 0  # subsurface = 0
    params[:, 11] = np.maximum(params[:, 11], 40)  # eta*0.5 away from 0 (eta == 0 is NaN in the reference too)
    p = params.reshape(n, 4, 4)
    packed = (p[..., 0] | (p[..., 1] << 8) | (p[..., 2] << 16) | (p[..., 3] << 24)).astype(np.uint32)
    return {"color": rng.uniform(0.02, 1.0, size=(n, 3)).astype(np.float32), "absorption": rng.uniform(0, 0.5, size=(n, 3)).astype(np.float32),
            "params": packed, "N": N, "wo": wo, "wi": wi, "t": rng.uniform(0.1, 5, size=n).astype(np.float32),
            "backfacing": rng.integers(0, 2, size=n).astype(np.int32), "r3": rng.uniform(0, 1, size=n).astype(np.float32),
            "r4": rng.uniform(0, 1, size=n).astype(np.float32)}


def patched_materials(sc):
    """DeviceMaterial array with texel offsets patched like CUDART/src/Context.cpp:167-191,201-268"""
    offs, pool, at = [], [], 0
    for t in sc.textures:
        offs.append(at)
        if t["type"] == R.TEX_UINT:
            pool.append(np.asarray(t["data"], np.uint32))
            at += t["data"].size
    m = sc.materials.copy()
    for i in range(len(m)):
        for slot, name in ((0, "tex0"), (1, "tex1"), (2, "tex2"), (3, "nmap0"), (4, "nmap1"), (5, "nmap2")):
            tid = sc.tex_ids[i, slot]
            if tid >= 0:
                m[name]["texaddr"][i] = offs[tid]
    return m, (np.concatenate(pool) if pool else np.zeros(4, np.uint32))


def soup_scene():
    sc = S.feature_soup(800, seed=21)
    S.extract_area_lights(sc)
    return sc


def soup_inputs(rng, ref=None):
    """inputs for the traversal / shading-data / light checks on the feature soup; with `ref` also the reference outputs"""
    sc = soup_scene()
    out = {}
    # traversal on mesh 0 (indexed) and mesh 1 (unindexed): rays through the cluster
    n = 400
    o = (rng.normal(size=(n, 3)) * 2.0).astype(np.float32)
    tgt = rng.uniform(-1, 1, size=(n, 3)).astype(np.float32)
    d = tgt - o
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    out["trav_o"], out["trav_d"] = o, d
    out["trav_tmax"] = rng.uniform(0.5, 4.0, size=n).astype(np.float32)
    # shading data queries: (instance, prim, u, v, D, cone)
    q = 300
    inst = rng.integers(0, 5, size=q).astype(np.int32)
    prim = np.array([rng.integers(0, len(sc.meshes[sc.instances[i][0]].triangles)) for i in inst], np.int32)
    u = rng.uniform(0, 1, size=q)
    v = rng.uniform(0, 1, size=q) * (1 - u)
    D = rng.normal(size=(q, 3))
    D = (D / np.linalg.norm(D, axis=1, keepdims=True)).astype(np.float32)
    out["sd_inst"], out["sd_prim"], out["sd_u"], out["sd_v"], out["sd_D"] = inst, prim, u.astype(np.float32), v.astype(np.float32), D
    out["sd_cone"] = rng.uniform(1e-4, 5e-2, size=q).astype(np.float32)
    # light queries
    k = 300
    out["li_I"] = rng.uniform(-2, 2, size=(k, 3)).astype(np.float32)
    Nn = rng.normal(size=(k, 3))
    out["li_N"] = (Nn / np.linalg.norm(Nn, axis=1, keepdims=True)).astype(np.float32)
    out["li_r"] = rng.uniform(0, 1, size=(k, 2)).astype(np.float32)
    out["li_O"] = rng.uniform(-2, 2, size=(k, 3)).astype(np.float32)
    if ref is not None:
        out.update(reference_soup_outputs(ref, sc, out))
    return out


def reference_soup_outputs(ref, sc, inp):
    """run the soup queries through the reference headers"""
    res = {}
    orc = R.RenderContext(R.load_oracle())
    S.upload(orc, sc, 16, 16)
    trav = ref.f("traverse_mbvh", I, [P, P, P, P, P, P, F, P, P, P])
    occl = ref.f("occluded_mbvh", I, [P, P, P, P, P, P, F, F])
    exp = orc.L.fn("export_mesh_mbvh", C.c_int, [P, C.c_size_t, P, C.c_size_t, P, C.c_size_t, P, P])
    for mi in (0, 1):
        nn, npm = C.c_size_t(), C.c_size_t()
        exp(orc._h, mi, None, 0, None, 0, C.byref(nn), C.byref(npm))
        nodes = np.zeros(nn.value * 32, np.float32)
        prims = np.zeros(npm.value, np.uint32)
        exp(orc._h, mi, nodes.ctypes.data, nn.value, prims.ctypes.data, npm.value, C.byref(nn), C.byref(npm))
        mesh = sc.meshes[mi]
        verts = np.ascontiguousarray(mesh.vertices, np.float32)
        idx = None if mesh.indices is None else np.ascontiguousarray(mesh.indices, np.uint32)
        rows = []
        for r in range(len(inp["trav_o"])):
            t, prim, b = F(1e34), I(-1), np.zeros(2, np.float32)
            hit = trav(nodes.ctypes.data, prims.ctypes.data, verts.ctypes.data, None if idx is None else idx.ctypes.data, fp(inp["trav_o"][r]),
                       fp(inp["trav_d"][r]), 1e-5, C.byref(t), C.byref(prim), b.ctypes.data)
            oc = occl(nodes.ctypes.data, prims.ctypes.data, verts.ctypes.data, None if idx is None else idx.ctypes.data, fp(inp["trav_o"][r]),
                      fp(inp["trav_d"][r]), 1e-5, float(inp["trav_tmax"][r]))
            rows.append((hit, t.value if hit else 1e34, prim.value if hit else -1, oc))
        res[f"trav_out_mesh{mi}"] = np.array(rows, np.float64)
    # getShadingData
    mats, pool = patched_materials(sc)
    gsd = ref.f("get_shading_data", None, [P, P, P, P, F, F, F, P, P, P, P, P])
    rows = []
    for i in range(len(inp["sd_inst"])):
        mesh_idx, M = sc.instances[int(inp["sd_inst"][i])]
        tri = sc.meshes[mesh_idx].triangles[int(inp["sd_prim"][i]):int(inp["sd_prim"][i]) + 1].copy()
        nm = np.linalg.inv(np.asarray(M)[:3, :3]).T
        nm9 = np.ascontiguousarray(nm.T.reshape(-1), np.float32)  # column-major
        uo, vo = float(inp["sd_u"][i]), float(inp["sd_v"][i])
        color, flags, N, iN = np.zeros(3, np.float32), U(), np.zeros(3, np.float32), np.zeros(3, np.float32)
        # reference (CUDART) convention: u, v = weights of vertex0, vertex1  <=>  (1-u-v, u) of ours
        gsd(mats.ctypes.data, pool.ctypes.data, tri.ctypes.data, fp(inp["sd_D"][i]), np.float32(1.0 - uo - vo), np.float32(uo), float(inp["sd_cone"][i]),
            nm9.ctypes.data, color.ctypes.data, C.addressof(flags), N.ctypes.data, iN.ctypes.data)
        rows.append(np.concatenate([color, [flags.value], N, iN]))
    res["sd_out"] = np.array(rows, np.float32)
    # lights
    ref.f("set_lights", None, [U, P, U, P, U, P, U, P])(len(sc.area_lights), sc.area_lights.ctypes.data, len(sc.point_lights), sc.point_lights.ctypes.data,
                                                         len(sc.spot_lights), sc.spot_lights.ctypes.data, len(sc.dir_lights), sc.dir_lights.ctypes.data)
    rpl = ref.f("random_point_on_light", None, [F, F, P, P, P, P, P, P])
    lpp = ref.f("light_pick_prob", F, [I, P, P, P])
    rows = []
    for i in range(len(inp["li_I"])):
        Pp, pick, pdf, col = np.zeros(3, np.float32), F(), F(), np.zeros(3, np.float32)
        rpl(float(inp["li_r"][i, 0]), float(inp["li_r"][i, 1]), fp(inp["li_I"][i]), fp(inp["li_N"][i]), Pp.ctypes.data, C.addressof(pick), C.addressof(pdf), col.ctypes.data)
        pp = lpp(i % max(len(sc.area_lights), 1), fp(inp["li_O"][i]), fp(inp["li_N"][i]), fp(inp["li_I"][i]))
        rows.append(np.concatenate([Pp, [pick.value, pdf.value], col, [pp]]))
    res["li_out"] = np.array(rows, np.float32)
    res["rb_out"] = np.array([ref.random_barycentrics(float(r)) for r in inp["li_r"][:, 0]], np.float32)
    return res
