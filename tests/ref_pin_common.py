"""Shared helpers of tests/golden/make_ref_golden.py and tests/test_ref_pin.py: ctypes wrappers around
oracle/_ref/librfwref.so (the reference's headers) and around the matching oracle hooks, plus input generators."""
import ctypes as C
from pathlib import Path

import numpy as np

import rfwb200 as R
from oracle.oracle_lib import load_oracle
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
        self.L = load_oracle()
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
    orc = R.RenderContext(load_oracle())
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


# ---- the reference's own CUDA wavefront kernels compiled for the host (oracle/ref_build/ref_kernels_shim.cpp) --------------
REF_KERNELS_LIB = R.REPO_DIR / "oracle" / "_ref" / "librfwref_kernels.so"


class _RefMesh(C.Structure):
    _fields_ = [("vertices4", C.c_void_p), ("indices3", C.c_void_p), ("triangles160", C.c_void_p), ("mbvh_nodes", C.c_void_p),
                ("prim_indices", C.c_void_p)]


class _RefInstance(C.Structure):
    _fields_ = [("mesh", C.c_int), ("transform", C.c_float * 16), ("inverse", C.c_float * 16), ("normal", C.c_float * 16)]


class _RefScene(C.Structure):
    _fields_ = [("n_meshes", C.c_int), ("meshes", C.c_void_p), ("n_instances", C.c_int), ("instances", C.c_void_p),
                ("tlas_nodes", C.c_void_p), ("tlas_prims", C.c_void_p), ("materials192", C.c_void_p), ("uint_texels", C.c_void_p),
                ("float_texels4", C.c_void_p), ("sky3", C.c_void_p), ("sky_w", C.c_uint), ("sky_h", C.c_uint),
                ("n_area", C.c_uint), ("n_point", C.c_uint), ("n_spot", C.c_uint), ("n_dir", C.c_uint),
                ("area", C.c_void_p), ("point", C.c_void_p), ("spot", C.c_void_p), ("dir", C.c_void_p), ("blue_noise", C.c_void_p)]


def pin_scene(rich=False, lights=False):
    """A box that keeps the oracle's documented deviations from CUDART out of play (oracle header D1-D6).
    rich=False: ONE emissive triangle whose material is material 0 (D1: the reference reads the material index where the
    light-triangle index is meant — both are 0 here), no alpha (D2), untextured, flat normals.
    rich=True adds what the oracle setting cudart_conventions=on covers: a two-triangle light (material 0 < light count, so the
    reference's potential[material] is a defined value), an indexed, bumpy, diffuse- and normal-mapped floor with smooth
    normals (trilinear fetch, LOD, texture coordinates and normals from the reference's area-ratio barycentrics), a smooth
    textured column and a second, instanced and scaled copy of it."""
    s = S.Scene(name="pin-box-rich" if rich else "pin-box")
    light = S.add_material(s, (17, 12, 4))  # material 0 = the light
    white = S.add_material(s, (0.73, 0.73, 0.73))
    red = S.add_material(s, (0.65, 0.05, 0.05))
    green = S.add_material(s, (0.12, 0.45, 0.15))
    glossy = S.add_material(s, (0.8, 0.8, 0.85), roughness=0.25, metallic=0.6, specular=0.5)
    c = 2.775
    if rich:
        lmesh = S.quad((0, -1, 0), (2.78, 5.549, 2.6), 1.5, 1.2, light)
    else:
        lp = np.array([[[2.0, 5.549, 2.0], [3.5, 5.549, 2.2], [2.6, 5.549, 3.4]]], np.float32)
        ln = np.array([[[0, -1, 0]] * 3], np.float32)
        ltri = S.make_triangles(lp, ln, np.zeros((1, 3, 2), np.float32), light)
        lmesh = S.SceneMesh(np.concatenate([lp.reshape(3, 3), np.ones((3, 1), np.float32)], 1), ltri, None)
    floor = S.quad((0, 1, 0), (c, 0, c), 5.55, 5.55, white)
    if rich:
        t0 = S.add_texture_rgba8(s, S.checker_texture(64, 5))
        n0 = S.add_texture_rgba8(s, S.normal_texture(64, 6))
        t1 = S.add_texture_rgba8(s, S.checker_texture(32, 9))
        tiles = S.add_material(s, (0.9, 0.9, 0.9), roughness=0.7, tex0=t0, nmap0=n0, uvscale=(1.0, 1.0))
        column = S.add_material(s, (0.85, 0.8, 0.7), roughness=0.45, specular=0.3, tex0=t1)
        # a smooth dielectric (roughness below MIN_ROUGHNESS: the IS_SPECULAR branches — no NEE, emissive hits without MIS,
        # refraction through SampleBSDF's transmission lobe with absorption) and a clear-coated, tinted metal
        glass = S.add_material(s, (0.95, 0.97, 1.0), roughness=0.0, transmission=0.9, eta=1.45, specular=0.6, absorption=(0.3, 0.1, 0.05))
        coated = S.add_material(s, (0.7, 0.25, 0.1), roughness=0.35, metallic=0.8, clearcoat=0.9, clearcoat_gloss=0.7, spec_tint=0.5,
                                subsurface=0.2)
        floor = S._grid_quad((0, 0, 0), (0, 0, 5.55), (5.55, 0, 0), 6, 6, tiles, uv_rep=3.0, tex_dims=(64, 64), bump=0.06,
                             rng=np.random.default_rng(3))
    s.meshes = [
        lmesh,  # mesh 0: its triangles are the area lights
        floor, S.quad((0, -1, 0), (c, 5.55, c), 5.55, 5.55, white),
        S.quad((0, 0, -1), (c, c, 5.55), 5.55, 5.55, white), S.quad((1, 0, 0), (0, c, c), 5.55, 5.55, green),
        S.quad((-1, 0, 0), (5.55, c, c), 5.55, 5.55, red),
        S.box_mesh((-0.5, 0, -0.5), (0.5, 1, 0.5), glass if rich else white),
        S.box_mesh((-0.5, 0, -0.5), (0.5, 1, 0.5), coated if rich else glossy),
    ]
    I = np.eye(4)
    s.instances = [(i, I) for i in range(6)]
    s.instances.append((6, S.translate(1.85, 0, 1.69) @ S.rotate_y(-18) @ S.scale(1.65, 1.65, 1.65)))
    s.instances.append((7, S.translate(3.68, 0, 3.51) @ S.rotate_y(15) @ S.scale(1.65, 3.30, 1.65)))
    if rich:
        s.meshes.append(S._cylinder((0, 0, 0), 0.35, 2.2, 14, 6, column, tex_dims=(32, 32)))
        s.instances.append((8, S.translate(4.4, 0.02, 1.1)))
        s.instances.append((8, S.translate(1.0, 0.02, 3.9) @ S.rotate_y(40) @ S.scale(1.4, 0.7, 0.9)))
    s.sky = (np.full((1, 3), 0.05, np.float32), 1, 1)
    if lights:  # every light type of lights.h in one pick table, and a sky with texture (the lookup of Kernels.cu:593-600)
        import math

        pl = np.zeros(1, R.POINT_LIGHT_DTYPE)
        pl["position"], pl["radiance"] = (1.2, 4.2, 0.8), (5, 4, 3)
        pl["energy"] = np.linalg.norm(pl["radiance"][0])
        sl = np.zeros(1, R.SPOT_LIGHT_DTYPE)
        sl["position"], sl["radiance"], sl["direction"] = (4.6, 4.8, 0.6), (14, 14, 18), (-0.35, -0.8, 0.45)
        sl["direction"] /= np.linalg.norm(sl["direction"][0])
        sl["cos_inner"], sl["cos_outer"] = math.cos(math.radians(14)), math.cos(math.radians(32))
        sl["energy"] = np.linalg.norm(sl["radiance"][0])
        dl = np.zeros(1, R.DIR_LIGHT_DTYPE)
        dl["direction"], dl["radiance"] = (0.25, -0.6, 0.75), (0.9, 0.85, 0.7)
        dl["direction"] /= np.linalg.norm(dl["direction"][0])
        dl["energy"] = np.linalg.norm(dl["radiance"][0])
        s.point_lights, s.spot_lights, s.dir_lights = pl, sl, dl
        sw, sh = 32, 16
        yy, xx = np.mgrid[0:sh, 0:sw]
        s.sky = (np.stack([0.2 + 0.5 * xx / sw, 0.3 + 0.4 * yy / sh, 0.9 - 0.4 * yy / sh], -1).astype(np.float32).reshape(-1, 3), sw, sh)
    s.camera_pos, s.camera_dir, s.fov = (2.78, 2.73, -8.0), (0, 0, 1), 40.0
    return s


def reference_kernels_scene(orc_ctx, sc):
    """Marshal the scene `sc` (already uploaded to the oracle context `orc_ctx`, whose MBVHs are exported in the reference's node
    layout) for the reference's host-compiled kernels.  -> (RefScene struct, objects that own its memory)"""
    P = C.c_void_p
    exp_mesh = orc_ctx.L.fn("export_mesh_mbvh", C.c_int, [P, C.c_size_t, P, C.c_size_t, P, C.c_size_t, P, P])
    exp_tlas = orc_ctx.L.fn("export_tlas_mbvh", C.c_int, [P, P, C.c_size_t, P, C.c_size_t, P, P])
    keep = []

    def arr(a, dt):
        a = np.ascontiguousarray(a, dt)
        keep.append(a)
        return a.ctypes.data

    meshes = (_RefMesh * len(sc.meshes))()
    for mi, m in enumerate(sc.meshes):
        nn, npm = C.c_size_t(), C.c_size_t()
        exp_mesh(orc_ctx._h, mi, None, 0, None, 0, C.byref(nn), C.byref(npm))
        nodes, prims = np.zeros(nn.value * 32, np.float32), np.zeros(max(npm.value, 1), np.uint32)
        exp_mesh(orc_ctx._h, mi, nodes.ctypes.data, nn.value, prims.ctypes.data, npm.value, C.byref(nn), C.byref(npm))
        keep += [nodes, prims]
        meshes[mi].vertices4 = arr(m.vertices, np.float32)
        meshes[mi].indices3 = None if m.indices is None else arr(m.indices, np.uint32)
        meshes[mi].triangles160 = arr(m.triangles, R.TRIANGLE_DTYPE)
        meshes[mi].mbvh_nodes, meshes[mi].prim_indices = nodes.ctypes.data, prims.ctypes.data
    insts = (_RefInstance * len(sc.instances))()
    exp_inst = orc_ctx.L.fn("export_instance", C.c_int, [P, C.c_size_t, P, P, P])
    for ii, (mi, _) in enumerate(sc.instances):
        # the matrices the oracle itself uses (its float32 inverse restates glm::inverse, top_level_bvh.cpp:309): last-bit
        # differences in the inverse decide self-intersections at the reference's fixed 1e-5 epsilons
        t16, i16, n9 = np.zeros(16, np.float32), np.zeros(16, np.float32), np.zeros(9, np.float32)
        assert exp_inst(orc_ctx._h, ii, t16.ctypes.data, i16.ctypes.data, n9.ctypes.data) == 0
        n16 = np.eye(4, dtype=np.float32)
        n16[:3, :3] = n9.reshape(3, 3)  # both column-major: columns stay columns
        insts[ii].mesh = mi
        insts[ii].transform[:] = [float(x) for x in t16]
        insts[ii].inverse[:] = [float(x) for x in i16]
        insts[ii].normal[:] = [float(x) for x in n16.reshape(-1)]
    nn, npm = C.c_size_t(), C.c_size_t()
    exp_tlas(orc_ctx._h, None, 0, None, 0, C.byref(nn), C.byref(npm))
    tnodes, tprims = np.zeros(nn.value * 32, np.float32), np.zeros(max(npm.value, 1), np.uint32)
    exp_tlas(orc_ctx._h, tnodes.ctypes.data, nn.value, tprims.ctypes.data, npm.value, C.byref(nn), C.byref(npm))
    mats, pool = patched_materials(sc)
    rs = _RefScene()
    rs.n_meshes, rs.meshes, rs.n_instances, rs.instances = len(sc.meshes), C.addressof(meshes), len(sc.instances), C.addressof(insts)
    rs.tlas_nodes, rs.tlas_prims = tnodes.ctypes.data, tprims.ctypes.data
    rs.materials192 = arr(mats, R.MATERIAL_DTYPE)
    rs.uint_texels = arr(pool if len(pool) else np.zeros(4, np.uint32), np.uint32)
    rs.float_texels4 = arr(np.zeros(4, np.float32), np.float32)
    sky, sw, sh = sc.sky
    rs.sky3, rs.sky_w, rs.sky_h = arr(np.asarray(sky, np.float32).reshape(-1, 3), np.float32), sw, sh
    for name, lights, dt in (("area", sc.area_lights, R.AREA_LIGHT_DTYPE), ("point", sc.point_lights, R.POINT_LIGHT_DTYPE),
                             ("spot", sc.spot_lights, R.SPOT_LIGHT_DTYPE), ("dir", sc.dir_lights, R.DIR_LIGHT_DTYPE)):
        setattr(rs, "n_" + name, len(lights))
        setattr(rs, name, arr(lights if len(lights) else np.zeros(1, dt), dt))
    rs.blue_noise = arr(blue_noise_uint_table(), np.uint32)
    keep += [meshes, insts, tnodes, tprims]
    return rs, keep


def reference_kernels_run(rs, view14, w, h, first, count, clamp=10.0, want_rays=True):
    """Render samples [first, first + count) with the reference's host-compiled kernels.
    -> dict(acc (h,w,4), origins, directions, states (w*h,4) of sample `first`, counters (count, 8, 3))"""
    lib = C.CDLL(str(REF_KERNELS_LIB))
    P = C.c_void_p
    n = w * h
    acc = np.zeros((n, 4), np.float32)
    o, d, st = ((np.zeros((n, 4), np.float32) for _ in range(3)) if want_rays else (None, None, None))
    counters = np.zeros((count, 8, 3), np.uint32)
    f = lib.rfwref_cudart_render
    f.restype, f.argtypes = C.c_int, [P, P, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_float, P, P, P, P, P]
    v = np.ascontiguousarray(view14, np.float32)
    ptr = lambda a: None if a is None else a.ctypes.data  # noqa: E731
    rc = f(C.addressof(rs), v.ctypes.data, w, h, first, count, clamp, acc.ctypes.data, ptr(o), ptr(d), ptr(st), counters.ctypes.data)
    assert rc == 0
    return {"acc": acc.reshape(h, w, 4), "origins": o, "directions": d, "states": st, "counters": counters}


def reference_blit(acc, samples):
    """The reference's own blit_buffer (Kernels.cu:181-203) on an accumulator (h, w, 4) -> finalised RGBA32F image."""
    lib = C.CDLL(str(REF_KERNELS_LIB))
    acc = np.ascontiguousarray(acc, np.float32)
    h, w = acc.shape[:2]
    img = np.zeros_like(acc)
    f = lib.rfwref_cudart_blit
    f.restype, f.argtypes = C.c_int, [C.c_void_p, C.c_uint, C.c_uint, C.c_uint, C.c_void_p]
    assert f(acc.ctypes.data, w, h, samples, img.ctypes.data) == 0
    return img


def reference_kernels_render(orc_ctx, sc, view14, w, h, first, count, clamp=10.0):
    rs, keep = reference_kernels_scene(orc_ctx, sc)
    return reference_kernels_run(rs, view14, w, h, first, count, clamp)


def reference_kernels_timed(rs, view14, w, h, spp, procs, timeout_s=240.0):
    """Time the reference's host-compiled kernels on `spp` samples of a w x h frame, the samples dealt round-robin to `procs`
    forked processes (the kernels keep their state in globals, so a process is the unit of parallelism; samples of a frame are
    independent given the sample index).  -> (seconds = slowest process, mean radiance of the frame as a sanity value)"""
    import os
    import struct
    import time

    procs = max(1, min(procs, spp))
    pipes = []
    for k in range(procs):
        r, wfd = os.pipe()
        pid = os.fork()
        if pid == 0:
            code = 1
            try:
                os.close(r)
                t0, total = time.perf_counter(), 0.0
                for s_ in range(k, spp, procs):
                    total += float(reference_kernels_run(rs, view14, w, h, s_, 1, want_rays=False)["acc"][..., :3].sum(dtype=np.float64))
                os.write(wfd, struct.pack("dd", time.perf_counter() - t0, total))
                code = 0
            finally:
                os._exit(code)
        os.close(wfd)
        pipes.append((pid, r))
    import select
    import signal

    secs, total, deadline, failed = 0.0, 0.0, time.monotonic() + timeout_s, False
    for pid, r in pipes:
        buf = b""
        while len(buf) < 16 and not failed:
            ready, _, _ = select.select([r], [], [], max(0.0, deadline - time.monotonic()))
            if not ready:
                failed = True  # out of time: stop waiting, the workers are killed below (by their exact pids)
                break
            chunk = os.read(r, 16 - len(buf))
            if not chunk:
                break
            buf += chunk
        os.close(r)
        if failed:
            os.kill(pid, signal.SIGKILL)
        _, status = os.waitpid(pid, 0)
        if failed or status != 0 or len(buf) < 16:
            failed = True
            continue
        dt, tot = struct.unpack("dd", buf)
        secs, total = max(secs, dt), total + tot
    if failed:
        raise RuntimeError("a reference-kernel worker failed or timed out")
    return secs, total / (spp * w * h * 3)


def pin_cases():
    """name -> (width, height, first sample, sample count, lens aperture).  'lens' has a wide lens (the blade sampling of
    generatePrimaryRay) at a packet-aligned size; 'long' crosses sample 256, where shade_rays switches from the blue-noise table
    to RandomFloat(seed) for the light sample, at a size that is not a multiple of the 8x8 / 64 / 128 launch shapes."""
    return {"lens": (64, 48, 0, 4, 0.12), "long": (70, 50, 0, 260, 0.0), "rich": (72, 56, 0, 8, 0.05),
            "lights": (64, 40, 0, 258, 0.0)}


def pin_view14(sc, w, h, aperture):
    cam = sc.camera(w, h)
    cam.aperture, cam.focalDistance = np.float32(aperture), np.float32(10.5)
    v = cam.get_view()
    return np.array(list(v.pos) + list(v.p1) + list(v.p2) + list(v.p3) + [v.aperture, v.spread_angle], np.float32)


def view_from14(v14):
    view = R.CameraView()
    for i in range(3):
        view.pos[i], view.p1[i], view.p2[i], view.p3[i] = float(v14[i]), float(v14[3 + i]), float(v14[6 + i]), float(v14[9 + i])
    view.aperture, view.spread_angle = float(v14[12]), float(v14[13])
    return view


# ---- the reference's own E-mode material lookup (EmbreeRT/src/Context.cpp:417-476 via oracle/ref_build/ref_emat_shim.cpp) -------
REF_EMAT_LIB = R.REPO_DIR / "oracle" / "_ref" / "librfwref_emat.so"


class _RefTexture(C.Structure):
    _fields_ = [("type", C.c_int), ("width", C.c_uint), ("height", C.c_uint), ("data", C.c_void_p)]


def ref_emode_material(sc, inst, prim, u, v):
    """retrieve_material of the reference for (instance, primitive) of scene `sc` at Embree barycentrics (u, v).
    -> (color3, N3, iN3).  Materials keep the texture ID in texaddr0, as the CPU backend's do."""
    lib = C.CDLL(str(REF_EMAT_LIB))
    mesh_idx, M = sc.instances[inst]
    tri = sc.meshes[mesh_idx].triangles[prim:prim + 1].copy()
    mat = sc.materials[int(tri["material"][0]):int(tri["material"][0]) + 1].copy()
    keep, texs = [], (_RefTexture * max(len(sc.textures), 1))()
    for i, t in enumerate(sc.textures):
        data = np.ascontiguousarray(t["data"])
        keep.append(data)
        texs[i].type, texs[i].width, texs[i].height, texs[i].data = (1 if t["type"] == R.TEX_UINT else 0), t["width"], t["height"], data.ctypes.data
    nm = np.eye(4)
    nm[:3, :3] = np.linalg.inv(np.asarray(M, np.float64)[:3, :3]).T
    nm16 = np.ascontiguousarray(nm.T.reshape(-1), np.float32)
    bary = np.array([1.0 - u - v, u, v], np.float32)
    color, N, iN = (np.zeros(3, np.float32) for _ in range(3))
    P = C.c_void_p
    f = lib.rfwref_emode_material
    f.restype, f.argtypes = None, [P, P, P, C.c_int, P, P, P, P, P]
    f(tri.ctypes.data, mat.ctypes.data, C.addressof(texs), len(sc.textures), bary.ctypes.data, nm16.ctypes.data, color.ctypes.data,
      N.ctypes.data, iN.ctypes.data)
    return color, N, iN


def emat_scene():
    """feature soup + a quad whose diffuse map is a FLOAT4 texture: retrieve_material's FLOAT4 case falls through into the
    UINT case (EmbreeRT/src/Context.cpp:458-472, no `break`), which the oracle restates."""
    sc = soup_scene()
    rng = np.random.default_rng(12)
    tf = S.add_texture_float4(sc, rng.uniform(0.2, 1.0, (16, 16, 4)).astype(np.float32))
    m = S.add_material(sc, (0.9, 0.8, 0.7), tex0=tf, uvscale=(2.0, 1.5), uvoffs=(0.25, -0.5))
    sc.meshes.append(S._grid_quad((0, 0, 0), (1, 0, 0), (0, 1, 0), 2, 2, m, uv_rep=1.0, tex_dims=(16, 16)))
    sc.instances.append((len(sc.meshes) - 1, S.translate(0.5, 0.5, -3.0) @ S.rotate_y(20) @ S.scale(1.5, 0.7, 1.0)))
    return sc


# ---- the reference's own E-mode frame loop body (EmbreeRT/src/Context.cpp:179-282 via oracle/ref_build/ref_eframe_shim.cpp) ------
REF_EFRAME_LIB = R.REPO_DIR / "oracle" / "_ref" / "librfwref_eframe.so"


class _RTCRay8(C.Structure):
    _fields_ = [(n, C.c_float * 8) for n in ("org_x", "org_y", "org_z", "tnear", "dir_x", "dir_y", "dir_z", "time", "tfar")] + \
               [("mask", C.c_uint * 8), ("id", C.c_int * 8), ("flags", C.c_uint * 8)]


class _RTCHit8(C.Structure):
    _fields_ = [(n, C.c_float * 8) for n in ("Ng_x", "Ng_y", "Ng_z", "u", "v")] + \
               [("primID", C.c_uint * 8), ("geomID", C.c_uint * 8), ("instID", (C.c_uint * 8) * 1)]


class _RTCRayHit8(C.Structure):
    _fields_ = [("ray", _RTCRay8), ("hit", _RTCHit8)]


class _RefEScene(C.Structure):
    _fields_ = [("materials192", C.c_void_p), ("n_materials", C.c_int), ("textures", C.c_void_p), ("n_textures", C.c_int),
                ("mesh_triangles160", C.c_void_p), ("n_meshes", C.c_int), ("instance_mesh", C.c_void_p), ("instance_normal16", C.c_void_p),
                ("n_instances", C.c_int), ("area_lights96", C.c_void_p), ("n_area", C.c_int), ("point_lights32", C.c_void_p), ("n_point", C.c_int),
                ("sky3", C.c_void_p), ("sky_w", C.c_int), ("sky_h", C.c_int)]


_OCCLUDED_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_float, C.c_float)


def ref_emode_frame(sc, o, W, H, probe=None, sample_index=0):
    """One E-mode frame of scene `sc` the way the reference's CPU renderer computes it AROUND Embree: the camera rays of the
    oracle context `o` (E-mode determinism contract) and their closest hits (the oracle's traversal in Embree's place) are put
    into Embree's ray / hit packets, eight pixels at a time, and handed to the reference's own per-pixel loop body
    (Context.cpp:179-282) — sky lookup, probe, retrieve_material, light loops, pixel write.  rtcOccluded1 is answered by the
    oracle's any-hit query with the oracle's stated shadow interval (t_min = the ray's tnear, t_max = tfar * (1 - 1e-4): a
    light's own triangle cannot shadow its centroid; DESIGN.md "E-mode determinism contract").
    -> (pixels (H, W, 4) float32, probe (instance, triangle, distance) or None)"""
    lib = C.CDLL(str(REF_EFRAME_LIB))
    S.extract_area_lights(sc)
    cam = sc.camera(W, H)
    o.set_setting("mode", "embree")
    origins, dirs = o.generate_primary(cam, sample_index)
    hits = o.trace_closest(origins, dirs)
    keep = []
    texs = (_RefTexture * max(len(sc.textures), 1))()
    for i, t in enumerate(sc.textures):
        data = np.ascontiguousarray(t["data"])
        keep.append(data)
        texs[i].type, texs[i].width, texs[i].height, texs[i].data = (1 if t["type"] == R.TEX_UINT else 0), t["width"], t["height"], data.ctypes.data
    mats = np.ascontiguousarray(sc.materials)
    tri_arrays = [np.ascontiguousarray(m.triangles) for m in sc.meshes]
    tri_ptrs = (C.c_void_p * len(tri_arrays))(*[a.ctypes.data for a in tri_arrays])
    inst_mesh = np.array([mi for mi, _ in sc.instances], np.uint32)
    nms = []
    for _, M in sc.instances:
        nm = np.eye(4)
        nm[:3, :3] = np.linalg.inv(np.asarray(M, np.float64)[:3, :3]).T
        nms.append(nm.T.reshape(-1))
    nm16 = np.ascontiguousarray(np.array(nms), np.float32)
    area = np.ascontiguousarray(sc.area_lights) if sc.area_lights is not None else np.zeros(0, R.AREA_LIGHT_DTYPE)
    point = np.ascontiguousarray(sc.point_lights) if sc.point_lights is not None else np.zeros(0, R.POINT_LIGHT_DTYPE)
    sky, sw, sh = sc.sky
    sky = np.ascontiguousarray(sky, np.float32)
    es = _RefEScene(mats.ctypes.data, len(mats), C.addressof(texs), len(sc.textures), C.addressof(tri_ptrs), len(tri_arrays), inst_mesh.ctypes.data,
                    nm16.ctypes.data, len(inst_mesh), area.ctypes.data if len(area) else None, len(area), point.ctypes.data if len(point) else None,
                    len(point), sky.ctypes.data, int(sw), int(sh))

    def occluded(_user, org, dr, tnear, tfar):
        oo = np.array([[org[0], org[1], org[2], 0.0]], np.float32)
        dd = np.array([[dr[0], dr[1], dr[2], 0.0]], np.float32)
        return int(o.trace_occluded(oo, dd, np.array([tfar * (1.0 - 1e-4)], np.float32), t_min=float(tnear))[0])

    cb = _OCCLUDED_FN(occluded)
    f = lib.rfwref_emode_shade_packet
    f.restype, f.argtypes = None, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, _OCCLUDED_FN, C.c_void_p, C.c_void_p, C.c_void_p]
    pixels = np.zeros((W * H, 4), np.float32)
    probe_out = np.array([0.0, 0.0, -1.0], np.float32)
    probe_id = probe[1] * W + probe[0] if probe is not None else -1
    n = W * H
    for base in range(0, n, 8):
        pk = _RTCRayHit8()
        for j in range(8):
            i = min(base + j, n - 1)
            pk.ray.org_x[j], pk.ray.org_y[j], pk.ray.org_z[j] = origins[i, 0], origins[i, 1], origins[i, 2]
            pk.ray.dir_x[j], pk.ray.dir_y[j], pk.ray.dir_z[j] = dirs[i, 0], dirs[i, 1], dirs[i, 2]
            pk.ray.tnear[j], pk.ray.tfar[j] = 1e-5, hits["t"][i]
            pk.ray.id[j] = base + j if base + j < n else n  # >= maxPixelID: skipped by the loop body
            hit = hits["prim_id"][i] >= 0
            pk.hit.geomID[j] = 0 if hit else 0xFFFFFFFF
            pk.hit.instID[0][j] = int(hits["inst_id"][i]) if hit else 0xFFFFFFFF
            pk.hit.primID[j] = int(hits["prim_id"][i]) if hit else 0xFFFFFFFF
            pk.hit.u[j], pk.hit.v[j] = hits["u"][i], hits["v"][i]
        f(C.addressof(es), C.addressof(pk), n, probe_id, cb, None, pixels.ctypes.data, probe_out.ctypes.data)
    got_probe = (int(probe_out[0]), int(probe_out[1]), float(probe_out[2])) if probe is not None else None
    return pixels.reshape(H, W, 4), got_probe


def _eframe_lights_scene():
    """the feature soup with the FLOAT4-textured quad, under a textured sky and with a second point light: every branch of the
    frame loop body (sky lookup, emissive early-out, both light loops, occluded and unoccluded shadow rays)"""
    sc = emat_scene()
    rng = np.random.default_rng(3)
    sc.sky = (rng.uniform(0.1, 2.0, (32 * 16, 3)).astype(np.float32), 32, 16)
    return sc


# name -> (scene, width, height, probe pixel); widths / heights are multiples of the reference's 4x2 packet tiles
EFRAME_CASES = {
    "soup": (_eframe_lights_scene, 96, 64, (48, 40)),
    "cornell": (lambda: S.cornell_box(unit_scale=True), 64, 48, (20, 30)),
}
