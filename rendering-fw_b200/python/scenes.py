"""Scene producers and the upload sequence of the reference's RenderSystem, for tests and bench.

These play the role of the *callers* of the plugin boundary (rfw::system + loaders), which SURVEY.md
§8 marks out of scope as product code; they only exist so the hot path can be fed the reference's
wire formats:
  * upload()            = rfw::system::synchronize call order (RFW/system/src/rfw/system.cpp:247-433)
  * extract_area_lights = system::update_area_lights (system.cpp:967-1032; applied with the
                          mathematically intended column-vector transform)
  * make_material       = HostMaterial::convertToDeviceMaterial (material_list.cpp:318-480)
  * build_mips          = texture::construct_mipmaps (texture.cpp:163-209)
  * quad()              = geometry::Quad (geometry/quad.cpp:6-42)
  * triangle LOD        = assimp/object.cpp:727-731
"""
from __future__ import annotations

import math
import struct
import zlib
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

import rfwb200 as R

MIPLEVELCOUNT = 5


# ------------------------------------------------------------------------------------------------
@dataclass
class SceneMesh:
    vertices: np.ndarray  # (nv, 4) float32
    triangles: np.ndarray  # TRIANGLE_DTYPE
    indices: np.ndarray | None = None  # (nt, 3) uint32


@dataclass
class Scene:
    name: str = "scene"
    meshes: list = field(default_factory=list)
    instances: list = field(default_factory=list)  # (mesh_idx, 4x4 float64)
    materials: np.ndarray = field(default_factory=lambda: np.zeros(0, R.MATERIAL_DTYPE))
    tex_ids: np.ndarray = field(default_factory=lambda: np.zeros((0, 11), np.int32))
    textures: list = field(default_factory=list)
    point_lights: np.ndarray = field(default_factory=lambda: np.zeros(0, R.POINT_LIGHT_DTYPE))
    spot_lights: np.ndarray = field(default_factory=lambda: np.zeros(0, R.SPOT_LIGHT_DTYPE))
    dir_lights: np.ndarray = field(default_factory=lambda: np.zeros(0, R.DIR_LIGHT_DTYPE))
    area_lights: np.ndarray = field(default_factory=lambda: np.zeros(0, R.AREA_LIGHT_DTYPE))
    sky: tuple = (np.zeros((1, 3), np.float32), 1, 1)
    sky_rgbe: np.ndarray | None = None  # (h, w, 4) uint8: the Radiance RGBE texels `sky` was decoded from (kept for the bake)
    camera_pos: tuple = (0, 0, 0)
    camera_dir: tuple = (0, 0, 1)
    fov: float = 40.0

    def camera(self, width, height) -> R.Camera:
        return R.Camera(self.camera_pos, self.camera_dir, self.fov, width, height)

    def triangle_count(self) -> int:
        return sum(len(self.meshes[m].triangles) for m, _ in self.instances)


# ------------------------------------------------------------------------------------------------
# materials / textures
# ------------------------------------------------------------------------------------------------
def add_material(scene: Scene, color, roughness=1.0, metallic=0.0, specular=0.0, transmission=0.0, eta=1.0,
                 subsurface=0.0, clearcoat=0.0, clearcoat_gloss=0.0, spec_tint=0.0, absorption=(0, 0, 0),
                 tex0=-1, nmap0=-1, smooth=True, has_alpha=False, uvscale=(1.0, 1.0), uvoffs=(0.0, 0.0)) -> int:
    m = np.zeros(1, R.MATERIAL_DTYPE)
    ids = np.full((1, 11), -1, np.int32)
    m["diffuse"][0] = np.asarray(color, np.float16)
    m["transmittance"][0] = np.asarray(absorption, np.float16)

    def ch(x):
        return int(np.float32(x) * np.float32(255.0)) & 255  # TOCHAR, material_list.cpp:316

    m["parameters"][0][0] = ch(metallic) | (ch(subsurface) << 8) | (ch(specular) << 16) | (ch(roughness) << 24)
    m["parameters"][0][1] = ch(spec_tint)
    m["parameters"][0][2] = ch(clearcoat) | (ch(clearcoat_gloss) << 8) | (ch(transmission) << 16) | (ch(eta * 0.5) << 24)
    flags = (1 if eta > 0 else 0)
    if tex0 >= 0:
        t = scene.textures[tex0]
        flags |= R.MAT_HAS_DIFFUSE_MAP
        m["tex0"]["width"][0], m["tex0"]["height"][0] = t["width"], t["height"]
        m["tex0"]["uscale"][0], m["tex0"]["vscale"][0] = uvscale
        m["tex0"]["uoffs"][0], m["tex0"]["voffs"][0] = uvoffs
        m["tex0"]["texaddr"][0] = tex0  # material_list writes the texture id; the backend patches the address
        ids[0, 0] = tex0
    if nmap0 >= 0:
        t = scene.textures[nmap0]
        flags |= R.MAT_HAS_NORMAL_MAP
        m["nmap0"]["width"][0], m["nmap0"]["height"][0] = t["width"], t["height"]
        m["nmap0"]["uscale"][0], m["nmap0"]["vscale"][0] = uvscale
        m["nmap0"]["uoffs"][0], m["nmap0"]["voffs"][0] = uvoffs
        m["nmap0"]["texaddr"][0] = nmap0
        ids[0, 3] = nmap0
    if smooth:
        flags |= R.MAT_SMOOTH_NORMALS
    if has_alpha:
        flags |= R.MAT_HAS_ALPHA
    m["flags"][0] = flags
    scene.materials = np.concatenate([scene.materials, m])
    scene.tex_ids = np.concatenate([scene.tex_ids, ids])
    return len(scene.materials) - 1


def material_rule(emissive, diffuse, transparent, opacity, shininess, shininess_strength, eta, reflectivity) -> dict:
    """What rfw::material_list::add(aiMaterial*) makes of the values assimp hands over (RFW/system/src/rfw/material_list.cpp:54-78;
    for an OBJ/MTL: Ke, Kd, Tf, d, Ns, -, Ni, -): emissive colour wins over diffuse, roughness = 1 - sqrt(min(Ns, 1024) / 1024),
    specular only from a shininess strength inside (0, 1), eta only when > 1, transmission = 1 - opacity when opacity != 0.
    Untouched fields keep the HostMaterial defaults (material_list.h:50-62).  Pinned on the reference's own lines by
    tests/test_ref_pin_ingest.py."""
    f = np.float32
    emissive, diffuse, transparent = (np.asarray(x, f) for x in (emissive, diffuse, transparent))
    m = {"metallic": f(0), "subsurface": f(0), "specular": f(0.5), "roughness": f(0.5), "eta": f(1), "transmission": f(0)}
    m["color"] = np.maximum(emissive if np.any(emissive != 0) else diffuse, f(0))
    m["absorption"] = np.maximum(transparent, f(0))
    m["metallic"] = max(f(reflectivity), f(0))
    if reflectivity > 0:
        m["metallic"] = f(reflectivity)
    m["specular"] = f(shininess_strength) if 0 < shininess_strength < 1 else f(0)
    m["roughness"] = max(f(0), f(1) - np.sqrt(min(f(shininess), f(1024)) / f(1024))) if shininess > 0 else f(1)
    if eta > 1:
        m["eta"] = f(eta)
    if opacity != 0:
        m["transmission"] = f(1) - max(f(opacity), f(0))
    return m


def build_mips(rgba8: np.ndarray) -> np.ndarray:
    """rgba8: (h, w) uint32 packed r|g<<8|b<<16|a<<24 -> 5 concatenated levels (texture.cpp:163-225)."""
    levels = [rgba8.astype(np.uint32)]
    for _ in range(1, MIPLEVELCOUNT):
        src = levels[-1]
        h, w = src.shape[0] >> 1, src.shape[1] >> 1
        if h == 0 or w == 0:
            levels.append(np.zeros((h, w), np.uint32))
            continue
        s = [src[0:2 * h:2, 0:2 * w:2], src[0:2 * h:2, 1:2 * w:2], src[1:2 * h:2, 0:2 * w:2], src[1:2 * h:2, 1:2 * w:2]]
        a = np.minimum(np.minimum(s[0] >> 24, s[1] >> 24), np.minimum(s[2] >> 24, s[3] >> 24))
        ch = []
        for sh in (0, 8, 16):
            ch.append((sum(((x >> sh) & 255) for x in s) >> 2).astype(np.uint32))
        levels.append((a << 24) + (ch[2] << 16) + (ch[1] << 8) + ch[0])
    return np.concatenate([l.reshape(-1) for l in levels]).astype(np.uint32)


def add_texture_rgba8(scene: Scene, rgba: np.ndarray) -> int:
    """rgba: (h, w, 4) uint8, row 0 = v origin."""
    h, w = rgba.shape[:2]
    p = rgba.astype(np.uint32)
    packed = p[..., 0] | (p[..., 1] << 8) | (p[..., 2] << 16) | (p[..., 3] << 24)
    scene.textures.append({"type": R.TEX_UINT, "width": w, "height": h, "data": build_mips(packed)})
    return len(scene.textures) - 1


def add_texture_float4(scene: Scene, rgba: np.ndarray) -> int:
    h, w = rgba.shape[:2]
    scene.textures.append({"type": R.TEX_FLOAT4, "width": w, "height": h, "data": np.ascontiguousarray(rgba, np.float32).reshape(-1)})
    return len(scene.textures) - 1


# ------------------------------------------------------------------------------------------------
# geometry helpers
# ------------------------------------------------------------------------------------------------
def make_triangles(positions: np.ndarray, normals: np.ndarray | None, uvs: np.ndarray | None, material,
                   tex_dims=None) -> np.ndarray:
    """positions/normals: (nt, 3, 3); uvs: (nt, 3, 2); material: int or (nt,) -> TRIANGLE_DTYPE[nt]"""
    p = np.asarray(positions, np.float32)
    nt = len(p)
    t = np.zeros(nt, R.TRIANGLE_DTYPE)
    e1, e2 = p[:, 1] - p[:, 0], p[:, 2] - p[:, 0]
    cr = np.cross(e1, e2).astype(np.float32)
    ln = np.linalg.norm(cr, axis=1).astype(np.float32)
    N = cr / np.maximum(ln, np.float32(1e-30))[:, None]
    if normals is None:
        normals = np.repeat(N[:, None, :], 3, axis=1)
    normals = np.asarray(normals, np.float32)
    # flip when inconsistent with all three vertex normals (assimp/object.cpp:649-651)
    flip = (np.einsum("ij,ij->i", N, normals[:, 0]) < 0) & (np.einsum("ij,ij->i", N, normals[:, 1]) < 0) & \
           (np.einsum("ij,ij->i", N, normals[:, 2]) < 0)
    N = np.where(flip[:, None], -N, N)
    t["vN0"], t["vN1"], t["vN2"] = normals[:, 0], normals[:, 1], normals[:, 2]
    t["Nx"], t["Ny"], t["Nz"] = N[:, 0], N[:, 1], N[:, 2]
    t["vertex0"], t["vertex1"], t["vertex2"] = p[:, 0], p[:, 1], p[:, 2]
    t["dummy1"] = t["dummy2"] = t["dummy3"] = 1.0
    t["area"] = 0.5 * ln
    t["light_tri_idx"] = -1
    t["material"] = material
    if uvs is not None:
        uvs = np.asarray(uvs, np.float32)
        t["u"] = uvs[:, :, 0]
        t["v"] = uvs[:, :, 1]
        if tex_dims is not None:
            tw, th = tex_dims
            ta = np.float32(tw * th) * np.abs((uvs[:, 1, 0] - uvs[:, 0, 0]) * (uvs[:, 2, 1] - uvs[:, 0, 1]) -
                                              (uvs[:, 2, 0] - uvs[:, 0, 0]) * (uvs[:, 1, 1] - uvs[:, 0, 1]))
            with np.errstate(divide="ignore", invalid="ignore"):
                lod = np.sqrt(np.maximum(0.5 * np.log2(ta / np.maximum(ln, 1e-30)), 0.0))
            t["LOD"] = np.nan_to_num(lod, nan=0.0, posinf=0.0, neginf=0.0).astype(np.float32)
    return t


def quad(N, pos, width, height, material) -> SceneMesh:
    """geometry::Quad (geometry/quad.cpp:6-42): two unindexed triangles."""
    N = np.asarray(N, np.float32)
    pos = np.asarray(pos, np.float32)
    tmp = np.array([0, 1, 0], np.float32) if abs(N[0]) > 0.9 else np.array([1, 0, 0], np.float32)
    T = np.cross(N, tmp)
    T = 0.5 * width * T / np.linalg.norm(T)
    Tn = T / np.linalg.norm(T)
    B = np.cross(Tn, N)
    B = 0.5 * height * B / np.linalg.norm(B)
    v = np.array([pos - B - T, pos + B - T, pos - B + T, pos + B - T, pos + B + T, pos - B + T], np.float32)
    tri_pos = v.reshape(2, 3, 3)
    normals = np.broadcast_to(N, (2, 3, 3)).copy()
    t = make_triangles(tri_pos, normals, np.zeros((2, 3, 2), np.float32), material)
    t["Nx"], t["Ny"], t["Nz"] = N[0], N[1], N[2]
    verts = np.concatenate([v, np.ones((6, 1), np.float32)], axis=1)
    return SceneMesh(verts, t, None)


def box_mesh(lo, hi, material, uv_scale=1.0, tex_dims=None) -> SceneMesh:
    lo, hi = np.asarray(lo, np.float32), np.asarray(hi, np.float32)
    c = np.array([[lo[0], lo[1], lo[2]], [hi[0], lo[1], lo[2]], [hi[0], hi[1], lo[2]], [lo[0], hi[1], lo[2]],
                  [lo[0], lo[1], hi[2]], [hi[0], lo[1], hi[2]], [hi[0], hi[1], hi[2]], [lo[0], hi[1], hi[2]]], np.float32)
    faces = [(0, 3, 2, 1), (4, 5, 6, 7), (0, 1, 5, 4), (3, 7, 6, 2), (0, 4, 7, 3), (1, 2, 6, 5)]
    verts, idx, uvs, nrm = [], [], [], []
    for f in faces:
        base = len(verts)
        q = c[list(f)]
        n = np.cross(q[1] - q[0], q[2] - q[0])
        n = n / np.linalg.norm(n)
        for k, uv in zip(range(4), ((0, 0), (1, 0), (1, 1), (0, 1))):
            verts.append(q[k]), uvs.append(np.array(uv, np.float32) * uv_scale), nrm.append(n)
        idx += [(base, base + 1, base + 2), (base, base + 2, base + 3)]
    verts, uvs, nrm, idx = np.array(verts, np.float32), np.array(uvs, np.float32), np.array(nrm, np.float32), np.array(idx, np.uint32)
    t = make_triangles(verts[idx], nrm[idx], uvs[idx], material, tex_dims)
    return SceneMesh(np.concatenate([verts, np.ones((len(verts), 1), np.float32)], axis=1), t, idx)


def translate(x, y, z):
    m = np.eye(4)
    m[:3, 3] = (x, y, z)
    return m


def scale(sx, sy=None, sz=None):
    sy = sx if sy is None else sy
    sz = sx if sz is None else sz
    return np.diag([sx, sy, sz, 1.0])


def rotate_y(deg):
    a = math.radians(deg)
    m = np.eye(4)
    m[0, 0], m[0, 2], m[2, 0], m[2, 2] = math.cos(a), math.sin(a), -math.sin(a), math.cos(a)
    return m


# ------------------------------------------------------------------------------------------------
# system::update_area_lights (system.cpp:967-1032)
# ------------------------------------------------------------------------------------------------
def extract_area_lights(scene: Scene):
    """Caller-side stand-in for system::update_area_lights (system/src/rfw/system.cpp:985-1025): one AreaLight per emissive
    triangle of every instance, position = centroid, radiance = material colour, energy = |colour|.  Two choices differ from
    the reference for SCALED emissive instances (none of BASELINE's configs has one; the Sponza light is only translated):
    the normal is normalised and the area is the world-space area, where the reference keeps the unnormalised
    inverse-transpose product and the object-space `triangle.area` — inconsistent with its own world-space distances."""
    emissive = np.array([bool(np.any(m["diffuse"].astype(np.float32) > 1.0)) for m in scene.materials], bool)
    lights = []
    for inst_idx, (mesh_idx, M) in enumerate(scene.instances):
        mesh = scene.meshes[mesh_idx]
        tri = mesh.triangles
        idxs = np.nonzero(emissive[tri["material"]])[0] if len(emissive) else []
        M = np.asarray(M, np.float64)
        NM = np.linalg.inv(M[:3, :3]).T
        for ti in idxs:
            t = tri[ti]
            v = [(M[:3, :3] @ t[k].astype(np.float64) + M[:3, 3]).astype(np.float32) for k in ("vertex0", "vertex1", "vertex2")]
            n = NM @ np.array([t["Nx"], t["Ny"], t["Nz"]], np.float64)
            n = (n / np.linalg.norm(n)).astype(np.float32)
            color = scene.materials[t["material"]]["diffuse"].astype(np.float32)
            L = np.zeros(1, R.AREA_LIGHT_DTYPE)
            L["vertex0"], L["vertex1"], L["vertex2"] = v
            L["position"] = (v[0] + v[1] + v[2]) * np.float32(1.0 / 3.0)
            L["energy"] = np.float32(np.linalg.norm(color))
            L["radiance"] = color
            L["normal"] = n
            L["tri_idx"], L["inst_idx"] = ti, inst_idx
            L["area"] = np.float32(0.5 * np.linalg.norm(np.cross(v[1] - v[0], v[2] - v[0])))
            tri["light_tri_idx"][ti] = len(lights)
            tri["area"][ti] = L["area"][0]
            lights.append(L)
    scene.area_lights = np.concatenate(lights) if lights else np.zeros(0, R.AREA_LIGHT_DTYPE)


def upload(ctx: R.RenderContext, scene: Scene, width: int, height: int):
    """rfw::system::set_target + synchronize (system.cpp:198-223, 247-433)."""
    ctx.init(width, height)
    ctx.set_sky(*scene.sky)
    ctx.set_textures(scene.textures)
    ctx.set_materials(scene.materials, scene.tex_ids)
    extract_area_lights(scene)
    for i, m in enumerate(scene.meshes):
        ctx.set_mesh(i, m.vertices, m.triangles, m.indices)
    for i, (mesh_idx, M) in enumerate(scene.instances):
        ctx.set_instance(i, mesh_idx, M)
    ctx.set_lights(scene.area_lights, scene.point_lights, scene.spot_lights, scene.dir_lights)
    ctx.update()


# ------------------------------------------------------------------------------------------------
# config 1: Cornell box — 555-unit box, two blocks, emissive ceiling quad
# ------------------------------------------------------------------------------------------------
def cornell_box(unit_scale: bool = False) -> Scene:
    """unit_scale: instance the whole 555-unit box at scale 0.01 so the reference's fixed 1e-5 epsilons are
    meaningful (at 555 units one float ulp is 6e-5 and secondary rays self-intersect chaotically, in the
    reference as well — DESIGN.md 'epsilons')."""
    s = Scene(name="cornell-unit" if unit_scale else "cornell")
    white = add_material(s, (0.73, 0.73, 0.73))
    red = add_material(s, (0.65, 0.05, 0.05))
    green = add_material(s, (0.12, 0.45, 0.15))
    light = add_material(s, (17, 12, 4))
    glossy = add_material(s, (0.8, 0.8, 0.85), roughness=0.25, metallic=0.6, specular=0.5)
    c = 277.5
    s.meshes = [
        quad((0, 1, 0), (c, 0, c), 555, 555, white),  # floor
        quad((0, -1, 0), (c, 555, c), 555, 555, white),  # ceiling
        quad((0, 0, -1), (c, c, 555), 555, 555, white),  # back
        quad((1, 0, 0), (0, c, c), 555, 555, green),  # right wall (x = 0)
        quad((-1, 0, 0), (555, c, c), 555, 555, red),  # left wall (x = 555)
        quad((0, -1, 0), (c, 554.9, c), 130, 105, light),  # light
        box_mesh((-0.5, 0, -0.5), (0.5, 1, 0.5), white),  # unit block, instanced twice
        box_mesh((-0.5, 0, -0.5), (0.5, 1, 0.5), glossy),
    ]
    I = np.eye(4)
    s.instances = [(i, I) for i in range(6)]
    s.instances.append((6, translate(185, 0, 169) @ rotate_y(-18) @ scale(165, 165, 165)))  # short block
    s.instances.append((7, translate(368, 0, 351) @ rotate_y(15) @ scale(165, 330, 165)))  # tall block
    s.sky = (np.zeros((1, 3), np.float32), 1, 1)
    s.camera_pos, s.camera_dir, s.fov = (278, 273, -800), (0, 0, 1), 40.0
    if unit_scale:
        k = scale(0.01)
        s.instances = [(m, k @ M) for m, M in s.instances]
        s.camera_pos = (2.78, 2.73, -8.0)
    return s


# ------------------------------------------------------------------------------------------------
# feature soup: every branch of the path (textures + mips + alpha, normal map, all light types,
# specular / transmissive / subsurface materials, instancing with non-uniform transforms)
# ------------------------------------------------------------------------------------------------
def checker_texture(size, seed, alpha_holes=False):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:size, 0:size]
    c = ((x // max(size // 8, 1)) + (y // max(size // 8, 1))) & 1
    base = rng.integers(60, 255, size=(2, 3))
    img = np.zeros((size, size, 4), np.uint8)
    img[..., :3] = np.where(c[..., None] == 1, base[0], base[1])
    img[..., :3] = np.clip(img[..., :3].astype(np.int32) + rng.integers(-20, 20, size=(size, size, 3)), 0, 255)
    img[..., 3] = 255
    if alpha_holes:
        r = np.hypot(x - size / 2, y - size / 2)
        img[..., 3] = np.where(r < size / 4, 0, 255)
    return img


def normal_texture(size, seed):
    rng = np.random.default_rng(seed)
    n = rng.normal(0, 0.25, size=(size, size, 3))
    n[..., 2] = 1.0
    n /= np.linalg.norm(n, axis=2, keepdims=True)
    img = np.zeros((size, size, 4), np.uint8)
    img[..., :3] = np.clip((n * 0.5 + 0.5) * 255, 0, 255)
    img[..., 3] = 255
    return img


def feature_soup(n_tris=2000, seed=7) -> Scene:
    rng = np.random.default_rng(seed)
    s = Scene(name="soup")
    t0 = add_texture_rgba8(s, checker_texture(64, 1))
    t1 = add_texture_rgba8(s, checker_texture(128, 2, alpha_holes=True))
    tn = add_texture_rgba8(s, normal_texture(64, 3))
    mats = [
        add_material(s, (0.8, 0.7, 0.6)),
        add_material(s, (1.0, 1.0, 1.0), tex0=t0, uvscale=(2.0, 2.0)),
        add_material(s, (1.0, 1.0, 1.0), tex0=t1, has_alpha=True),
        add_material(s, (0.9, 0.9, 0.9), tex0=t0, nmap0=tn),
        add_material(s, (0.9, 0.9, 0.95), roughness=0.0, metallic=1.0),  # specular (roughness < 0.01)
        add_material(s, (0.7, 0.9, 0.7), roughness=0.3, transmission=0.8, eta=1.5, absorption=(0.1, 0.02, 0.1)),
        add_material(s, (0.8, 0.5, 0.4), roughness=0.6, subsurface=0.5, clearcoat=0.5, clearcoat_gloss=0.7, specular=0.4, spec_tint=0.3),
        add_material(s, (0.5, 0.5, 0.9), roughness=0.5, smooth=False),
    ]
    light_mat = add_material(s, (30, 28, 25))
    # random triangle clusters as a few meshes
    n_meshes = 4
    per = n_tris // n_meshes
    for mi in range(n_meshes):
        centers = rng.uniform(-1, 1, size=(per, 1, 3))
        pos = centers + rng.normal(0, 0.12, size=(per, 3, 3))
        nrm = np.cross(pos[:, 1] - pos[:, 0], pos[:, 2] - pos[:, 0])
        nrm /= np.maximum(np.linalg.norm(nrm, axis=1, keepdims=True), 1e-20)
        vn = nrm[:, None, :] + rng.normal(0, 0.15, size=(per, 3, 3))
        vn /= np.linalg.norm(vn, axis=2, keepdims=True)
        uv = rng.uniform(-1, 2, size=(per, 3, 2))
        mat = rng.choice(mats, size=per)
        tri = make_triangles(pos, vn, uv, mat.astype(np.uint32), tex_dims=(64, 64))
        if mi % 2 == 0:  # indexed with shared vertex array
            verts = pos.reshape(-1, 3).astype(np.float32)
            idx = np.arange(per * 3, dtype=np.uint32).reshape(per, 3)
            perm = rng.permutation(per * 3)
            inv = np.empty_like(perm)
            inv[perm] = np.arange(per * 3)
            verts = verts[perm]
            idx = inv[idx].astype(np.uint32)
            s.meshes.append(SceneMesh(np.concatenate([verts, np.ones((len(verts), 1), np.float32)], 1), tri, idx))
        else:
            verts = pos.reshape(-1, 3).astype(np.float32)
            s.meshes.append(SceneMesh(np.concatenate([verts, np.ones((len(verts), 1), np.float32)], 1), tri, None))
    s.meshes.append(box_mesh((-4, -2.2, -4), (4, -2, 4), mats[1], uv_scale=4.0, tex_dims=(64, 64)))  # floor slab
    s.meshes.append(quad((0, -1, 0), (0, 0, 0), 1.5, 1.5, light_mat))
    s.instances = [
        (0, np.eye(4)),
        (1, translate(1.5, 0.2, 0.5) @ rotate_y(30) @ scale(0.8, 1.3, 0.7)),
        (2, translate(-1.6, 0.0, 0.3) @ scale(0.9)),
        (3, translate(0.2, 0.5, 1.8) @ rotate_y(-50) @ scale(-0.7, 0.7, 0.7)),  # mirrored
        (0, translate(0.0, 0.3, -2.0) @ rotate_y(90) @ scale(0.6)),  # second instance of mesh 0
        (4, np.eye(4)),
        (5, translate(0.3, 3.0, 0.2)),
    ]
    pl = np.zeros(1, R.POINT_LIGHT_DTYPE)
    pl["position"], pl["radiance"] = (2.5, 2.0, -2.0), (6, 5, 4)
    pl["energy"] = np.linalg.norm(pl["radiance"][0])
    sl = np.zeros(1, R.SPOT_LIGHT_DTYPE)
    sl["position"], sl["radiance"], sl["direction"] = (-2.5, 2.5, -1.5), (9, 9, 12), (0.6, -0.7, 0.4)
    sl["direction"] /= np.linalg.norm(sl["direction"][0])
    sl["cos_inner"], sl["cos_outer"] = math.cos(math.radians(15)), math.cos(math.radians(35))
    sl["energy"] = np.linalg.norm(sl["radiance"][0])
    dl = np.zeros(1, R.DIR_LIGHT_DTYPE)
    dl["direction"], dl["radiance"] = (-0.3, -1, 0.2), (0.8, 0.8, 0.7)
    dl["direction"] /= np.linalg.norm(dl["direction"][0])
    dl["energy"] = np.linalg.norm(dl["radiance"][0])
    s.point_lights, s.spot_lights, s.dir_lights = pl, sl, dl
    sw, sh = 64, 32
    yy, xx = np.mgrid[0:sh, 0:sw]
    sky = np.stack([0.3 + 0.4 * xx / sw, 0.4 + 0.3 * yy / sh, 0.8 - 0.3 * yy / sh], axis=-1).astype(np.float32)
    s.sky = (sky.reshape(-1, 3), sw, sh)
    s.camera_pos, s.camera_dir, s.fov = (0.2, 0.6, -5.0), (0.0, -0.08, 1.0), 45.0
    return s


# ------------------------------------------------------------------------------------------------
# procedural atrium: stand-in for Sponza when the baked asset is not present (committed history has
# no binary assets).  A colonnaded two-storey hall with an open roof, textured walls, cloth banners
# and the config-2 light quad, tessellated to roughly `target_tris` triangles.
# ------------------------------------------------------------------------------------------------
def _grid_quad(p0, du, dv, nu, nv, material, uv_rep=1.0, tex_dims=None, bump=None, rng=None):
    """tessellated rectangle p0 + a*du + b*dv, optional normal-direction displacement."""
    p0, du, dv = (np.asarray(a, np.float64) for a in (p0, du, dv))
    a, b = np.meshgrid(np.linspace(0, 1, nu + 1), np.linspace(0, 1, nv + 1), indexing="xy")
    P = p0 + a[..., None] * du + b[..., None] * dv
    n = np.cross(du, dv)
    n = n / np.linalg.norm(n)
    if bump:
        disp = bump * np.sin(a * 9.0 + (rng.uniform(0, 6) if rng is not None else 0)) * np.cos(b * 5.0)
        P = P + disp[..., None] * n
    uv = np.stack([a * uv_rep, b * uv_rep], -1)
    idx = []
    W = nu + 1
    for j in range(nv):
        for i in range(nu):
            v00, v10, v01, v11 = j * W + i, j * W + i + 1, (j + 1) * W + i, (j + 1) * W + i + 1
            idx += [(v00, v10, v11), (v00, v11, v01)]
    idx = np.array(idx, np.uint32)
    verts = P.reshape(-1, 3).astype(np.float32)
    uvs = uv.reshape(-1, 2).astype(np.float32)
    # smooth normals from the displaced grid
    if bump:
        fn = np.cross(verts[idx[:, 1]] - verts[idx[:, 0]], verts[idx[:, 2]] - verts[idx[:, 0]])
        vn = np.zeros_like(verts)
        for k in range(3):
            np.add.at(vn, idx[:, k], fn)
        vn /= np.maximum(np.linalg.norm(vn, axis=1, keepdims=True), 1e-20)
    else:
        vn = np.broadcast_to(n.astype(np.float32), verts.shape).copy()
    tri = make_triangles(verts[idx], vn[idx], uvs[idx], material, tex_dims)
    return SceneMesh(np.concatenate([verts, np.ones((len(verts), 1), np.float32)], 1), tri, idx)


def _cylinder(base, radius, height, seg, rings, material, tex_dims=None):
    base = np.asarray(base, np.float64)
    th = np.linspace(0, 2 * np.pi, seg + 1)
    hh = np.linspace(0, 1, rings + 1)
    T, H = np.meshgrid(th, hh, indexing="xy")
    rad = radius * (1.0 + 0.08 * np.cos(H * np.pi * 2))
    P = np.stack([base[0] + rad * np.cos(T), base[1] + H * height, base[2] + rad * np.sin(T)], -1)
    Nn = np.stack([np.cos(T), np.zeros_like(T), np.sin(T)], -1)
    uv = np.stack([T / (2 * np.pi) * 2.0, H * 3.0], -1)
    W = seg + 1
    idx = []
    for j in range(rings):
        for i in range(seg):
            v00, v10, v01, v11 = j * W + i, j * W + i + 1, (j + 1) * W + i, (j + 1) * W + i + 1
            idx += [(v00, v11, v10), (v00, v01, v11)]
    idx = np.array(idx, np.uint32)
    verts, vn, uvs = P.reshape(-1, 3).astype(np.float32), Nn.reshape(-1, 3).astype(np.float32), uv.reshape(-1, 2).astype(np.float32)
    tri = make_triangles(verts[idx], vn[idx], uvs[idx], material, tex_dims)
    return SceneMesh(np.concatenate([verts, np.ones((len(verts), 1), np.float32)], 1), tri, idx)


# ------------------------------------------------------------------------------------------------
# config 4 ingredients: a skinned mesh (stand-in for CesiumMan when the asset is not baked): a tube with a chain of
# joints and smooth weights, standing on a floor under the config-2 light rule
# ------------------------------------------------------------------------------------------------
@dataclass
class Skin:
    mesh_index: int
    base_vertices: np.ndarray  # (nv, 4)
    base_normals: np.ndarray  # (nv, 3)
    joints: np.ndarray  # (nv, 4) uint32
    weights: np.ndarray  # (nv, 4)
    n_joints: int
    poses: object = None  # callable t -> (n_joints, 4, 4) float32 joint matrices, or an array (frames, n_joints, 4, 4)

    def joint_matrices(self, k: int) -> np.ndarray:
        if callable(self.poses):
            return self.poses(k)
        return np.asarray(self.poses[k % len(self.poses)], np.float32)


def _rot_z(a):
    m = np.eye(4)
    m[0, 0], m[0, 1], m[1, 0], m[1, 1] = math.cos(a), -math.sin(a), math.sin(a), math.cos(a)
    return m


def skinned_tube(seg=24, rings=32, n_joints=4, height=2.0, radius=0.25):
    """-> (Scene, Skin).  Mesh 0 = floor, 1 = light, 2 = the tube (indexed, smooth normals)."""
    s = Scene(name="skinned-tube")
    grey = add_material(s, (0.7, 0.7, 0.7))
    light = add_material(s, (20, 20, 18))
    skin_mat = add_material(s, (0.8, 0.45, 0.3), roughness=0.6)
    tube = _cylinder((0, 0, 0), radius, height, seg, rings, skin_mat)
    s.meshes = [quad((0, 1, 0), (0, 0, 0), 12, 12, grey), quad((0, -1, 0), (0, 4.5, 0), 3, 3, light), tube]
    s.instances = [(0, np.eye(4)), (1, np.eye(4)), (2, translate(0.2, 0.0, 0.1) @ rotate_y(20))]
    s.sky = (np.full((1, 3), 0.15, np.float32), 1, 1)
    s.camera_pos, s.camera_dir, s.fov = (0.0, 1.4, -5.0), (0.0, -0.05, 1.0), 40.0
    v = tube.vertices
    # joint j sits at height j * height / n_joints; a vertex is bound to the two (up to four) nearest joints
    seg_h = height / n_joints
    jpos = np.arange(n_joints) * seg_h
    d = np.abs(v[:, 1:2] - (jpos[None, :] + 0.5 * seg_h))
    order = np.argsort(d, axis=1)[:, :4]
    if order.shape[1] < 4:
        order = np.concatenate([order, np.repeat(order[:, :1], 4 - order.shape[1], 1)], 1)
    w = np.maximum(1.0 - np.take_along_axis(d, order, 1) / seg_h, 0.0)
    w[:, 2:] = 0.0
    w = (w / np.maximum(w.sum(1, keepdims=True), 1e-9)).astype(np.float32)
    nrm = np.zeros((len(v), 3), np.float32)
    idx = tube.indices
    for k, key in enumerate(("vN0", "vN1", "vN2")):
        nrm[idx[:, k]] = tube.triangles[key]

    def poses(k: int) -> np.ndarray:
        t = k / 60.0
        out, acc = [], np.eye(4)
        for j in range(n_joints):
            a = 0.35 * math.sin(2.0 * t + 0.9 * j)
            # rotate about the joint's own pivot (0, jpos[j], 0), accumulated down the chain
            acc = acc @ translate(0, jpos[j], 0) @ _rot_z(a) @ translate(0, -jpos[j], 0)
            out.append(acc.copy())
        return np.asarray(out, np.float32)

    return s, Skin(2, v.copy(), nrm, order.astype(np.uint32), w, n_joints, poses)


CESIUMMAN_BAKED = R.PKG_DIR / "data" / "_baked" / "cesiumman.npz"


def animated_config4(copies: int = 1):
    """BASELINE.json configs[3]: CesiumMan (baked by tools/bake_cesiumman.py; the skinned tube
    stands in when the bake is absent) on a 50x50 floor under the config-2 light quad (20x100, radiance 100, y = 60).
    -> (Scene, [Skin per skinned mesh]).  `copies` > 1 places several independently skinned copies (scaling runs)."""
    s = Scene(name="config4")
    grey = add_material(s, (0.6, 0.6, 0.6))
    light = add_material(s, (100, 100, 100))
    s.meshes = [quad((0, 1, 0), (0, 0, 0), 50, 50, grey), quad((0, -1, 0), (0, 0, 0), 20, 100, light)]
    s.instances = [(0, np.eye(4)), (1, translate(0, 60, 0))]
    skins = []
    if CESIUMMAN_BAKED.exists():
        d = np.load(CESIUMMAN_BAKED)
        s.name = "config4-cesiumman"
        tex = -1
        if d["texture"].size:
            tex = add_texture_rgba8(s, d["texture"])
        mat = add_material(s, (1, 1, 1), roughness=0.8, tex0=tex)
        pos, nrm, uv, idx = d["positions"], d["normals"], d["uvs"], d["indices"]
        tex_dims = d["texture"].shape[1::-1] if d["texture"].size else None
        tri = make_triangles(pos[idx], nrm[idx], uv[idx], mat, tex_dims)
        v4 = np.concatenate([pos, np.ones((len(pos), 1), np.float32)], 1)
        base_xf, poses = d["mesh_transform"], d["poses"]
        for c in range(copies):
            mi = len(s.meshes)
            s.meshes.append(SceneMesh(v4.copy(), tri.copy(), idx.copy()))
            s.instances.append((mi, translate(1.2 * ((c + 1) // 2) * (-1 if c % 2 else 1), 0, 0.8 * (c // 8)) @ base_xf))
            skins.append(Skin(mi, v4, nrm, d["joints"], d["weights"], poses.shape[1], np.roll(poses, 7 * c, axis=0)))
        s.camera_pos, s.camera_dir, s.fov = (0.6, 0.95, 2.3), (-0.25, -0.08, -1.0), 40.0
    else:
        tube_scene, sk = skinned_tube(seg=48, rings=48)
        mat = add_material(s, (0.8, 0.45, 0.3), roughness=0.6)
        tube = tube_scene.meshes[sk.mesh_index]
        tube.triangles["material"] = mat
        for c in range(copies):
            mi = len(s.meshes)
            s.meshes.append(SceneMesh(tube.vertices.copy(), tube.triangles.copy(), tube.indices.copy()))
            s.instances.append((mi, translate(1.2 * ((c + 1) // 2) * (-1 if c % 2 else 1), 0, 0.8 * (c // 8))))
            skins.append(Skin(mi, sk.base_vertices, sk.base_normals, sk.joints, sk.weights, sk.n_joints, sk.poses))
        s.camera_pos, s.camera_dir, s.fov = (0.0, 1.2, -5.0), (0.0, -0.04, 1.0), 40.0
    s.sky = (np.full((1, 3), 0.3, np.float32), 1, 1)
    return s, skins


def atrium(target_tris=262_000, seed=11) -> Scene:
    rng = np.random.default_rng(seed)
    s = Scene(name=f"atrium-{target_tris}")
    tex = [add_texture_rgba8(s, checker_texture(256, 10 + i)) for i in range(6)]
    cloth_tex = add_texture_rgba8(s, checker_texture(256, 30, alpha_holes=False))
    stone = [add_material(s, (1, 1, 1), roughness=0.9, tex0=tex[i], uvscale=(1, 1)) for i in range(6)]
    cloth = [add_material(s, c, roughness=0.8, tex0=cloth_tex) for c in ((0.8, 0.2, 0.2), (0.2, 0.7, 0.3), (0.2, 0.3, 0.8))]
    metal = add_material(s, (0.9, 0.8, 0.5), roughness=0.15, metallic=0.9, specular=0.6)
    light = add_material(s, (100, 100, 100), roughness=1.0)
    # Sponza-like proportions in the reference's scaled space (≈ 740 x 310 x 450 units at scale 0.2)
    L, Wd, Ht = 370.0, 160.0, 290.0  # half length (x), half width (z), height
    aisle = 55.0
    f = math.sqrt(target_tris / 148_000.0)  # the base tessellation below yields ~148k triangles
    k = lambda n: max(1, int(round(n * f)))
    dims = (256, 256)
    M = []
    M.append(_grid_quad((-L, 0, -Wd), (2 * L, 0, 0), (0, 0, 2 * Wd), k(120), k(56), stone[0], 12.0, dims, bump=0.3, rng=rng))  # floor
    M.append(_grid_quad((-L, 0, Wd), (2 * L, 0, 0), (0, Ht, 0), k(120), k(48), stone[1], 8.0, dims, bump=0.8, rng=rng))  # +z wall
    M.append(_grid_quad((L, 0, -Wd), (-2 * L, 0, 0), (0, Ht, 0), k(120), k(48), stone[1], 8.0, dims, bump=0.8, rng=rng))  # -z wall
    M.append(_grid_quad((-L, 0, -Wd), (0, 0, 2 * Wd), (0, Ht, 0), k(56), k(48), stone[2], 6.0, dims, bump=0.8, rng=rng))  # -x wall
    M.append(_grid_quad((L, 0, Wd), (0, 0, -2 * Wd), (0, Ht, 0), k(56), k(48), stone[2], 6.0, dims, bump=0.8, rng=rng))  # +x wall
    # upper gallery floors along both aisles + roof ring (open centre)
    for zsign in (-1, 1):
        z0 = zsign * aisle
        z1 = zsign * Wd
        M.append(_grid_quad((-L, 105, min(z0, z1)), (2 * L, 0, 0), (0, 0, abs(z1 - z0)), k(90), k(14), stone[3], 10.0, dims))
        M.append(_grid_quad((-L, 104, max(z0, z1)), (2 * L, 0, 0), (0, 0, -abs(z1 - z0)), k(90), k(14), stone[3], 10.0, dims))
        M.append(_grid_quad((-L, Ht, max(z0, z1)), (2 * L, 0, 0), (0, 0, -abs(z1 - z0)), k(60), k(10), stone[4], 10.0, dims))
    # columns: two storeys, both sides
    col = _cylinder((0, 0, 0), 9.0, 100.0, k(28), k(22), stone[5], dims)
    col_idx = len(M)
    M.append(col)
    ring = _cylinder((0, 0, 0), 11.0, 6.0, k(28), k(2), metal, dims)
    ring_idx = len(M)
    M.append(ring)
    # banners (cloth) hanging into the nave
    banner_idx = []
    for bi in range(3):
        banner_idx.append(len(M))
        M.append(_grid_quad((0, 0, 0), (0, 0, 40), (0, -90, 0), k(28), k(48), cloth[bi], 2.0, dims, bump=3.0, rng=rng))
    light_idx = len(M)
    M.append(quad((0, -1, 0), (0, 0, 0), 20.0, 100.0, light))
    s.meshes = M
    inst = [(i, np.eye(4)) for i in range(col_idx)]
    n_cols = 12
    for zsign in (-1, 1):
        for storey in (0, 1):
            for ci in range(n_cols):
                x = -L + (ci + 0.5) * (2 * L / n_cols)
                inst.append((col_idx, translate(x, 106 * storey, zsign * aisle) @ scale(1.0, 1.0 if storey == 0 else 0.9, 1.0)))
                inst.append((ring_idx, translate(x, 106 * storey + 2, zsign * aisle)))
    for bi, x in enumerate(np.linspace(-L * 0.7, L * 0.7, 9)):
        inst.append((banner_idx[bi % 3], translate(x, 200, -20)))
    inst.append((light_idx, translate(0, 260, 0) @ rotate_y(90)))
    s.instances = inst
    sw, sh = 256, 128
    yy, xx = np.mgrid[0:sh, 0:sw]
    sun = np.exp(-(((xx - 60) / 9.0) ** 2 + ((yy - 30) / 9.0) ** 2))
    sky = np.stack([0.5 + 3.0 * sun, 0.6 + 2.8 * sun, 0.8 + 2.0 * sun], axis=-1) * (1.0 - 0.5 * (yy / sh))[..., None]
    s.sky = (sky.astype(np.float32).reshape(-1, 3), sw, sh)
    s.camera_pos, s.camera_dir, s.fov = (-L * 0.9, 50.0, 0.0), (1.0, 0.0, 0.05), 40.0
    return s


# ------------------------------------------------------------------------------------------------
# baked scenes (tools/bake_sponza.py): a zlib-compressed container of the wire formats above
# ------------------------------------------------------------------------------------------------
BAKE_MAGIC = b"RFWB200S"


def decode_rgbe(rgbe: np.ndarray) -> np.ndarray:
    """Radiance RGBE texels -> float32 RGB the way FreeImage's HDR reader does for the reference's skybox
    (rfw::skybox::load, RFW/system/src/rfw/skybox.cpp:112-133, via FreeImage_ConvertToRGBAF): mantissa * 2^(e - 136), 0 when e == 0."""
    e = rgbe[..., 3].astype(np.int32)
    f = np.where(e > 0, np.ldexp(np.float32(1.0), e - 136), np.float32(0.0)).astype(np.float32)
    return (rgbe[..., :3].astype(np.float32) * f[..., None]).astype(np.float32)


def read_radiance_hdr(path) -> np.ndarray:
    """-> (h, w, 4) uint8 RGBE texels of a Radiance .hdr file (header '-Y h +X w', new-style run-length scanlines), top row first"""
    blob = Path(path).read_bytes()
    end = blob.index(b"\n\n") + 2
    nl = blob.index(b"\n", end)
    res = blob[end:nl].split()
    assert res[0] == b"-Y" and res[2] == b"+X", res
    h, w = int(res[1]), int(res[3])
    data = np.frombuffer(blob, np.uint8, offset=nl + 1)
    out = np.empty((h, w, 4), np.uint8)
    pos = 0
    for y in range(h):
        assert data[pos] == 2 and data[pos + 1] == 2 and (int(data[pos + 2]) << 8 | int(data[pos + 3])) == w, "only new-style RLE scanlines"
        pos += 4
        for c in range(4):
            x = 0
            while x < w:
                n = int(data[pos])
                if n > 128:
                    out[y, x:x + n - 128, c] = data[pos + 1]
                    x += n - 128
                    pos += 2
                else:
                    out[y, x:x + n, c] = data[pos + 1:pos + 1 + n]
                    x += n
                    pos += 1 + n
    return out


def save_baked(scene: Scene, path: Path):
    parts = []

    def put(arr):
        arr = np.ascontiguousarray(arr)
        parts.append(struct.pack("<Q", arr.nbytes))
        parts.append(arr.tobytes())

    hdr = struct.pack("<8sIIII", BAKE_MAGIC, len(scene.meshes), len(scene.instances), len(scene.materials), len(scene.textures))
    parts.append(hdr)
    name = scene.name.encode()
    parts.append(struct.pack("<I", len(name)) + name)
    parts.append(struct.pack("<7f", *scene.camera_pos, *scene.camera_dir, scene.fov))
    for m in scene.meshes:
        parts.append(struct.pack("<III", len(m.vertices), len(m.triangles), 0 if m.indices is None else 1))
        put(m.vertices.astype(np.float32)), put(m.triangles)
        if m.indices is not None:
            put(m.indices.astype(np.uint32))
    for mi, M in scene.instances:
        parts.append(struct.pack("<I", mi))
        put(np.asarray(M, np.float64))
    put(scene.materials), put(scene.tex_ids.astype(np.int32))
    for t in scene.textures:
        parts.append(struct.pack("<III", t["type"], t["width"], t["height"]))
        data = np.asarray(t["data"])
        if t["type"] == R.TEX_UINT:  # level 0 only: the mip chain is rebuilt by build_mips on load (keeps the snapshot small)
            level0 = data.reshape(-1)[: t["width"] * t["height"]]
            assert np.array_equal(build_mips(level0.reshape(t["height"], t["width"])), data.reshape(-1))
            data = level0
        put(data)
    sky, sw, sh = scene.sky
    if scene.sky_rgbe is not None:  # the asset's own RGBE texels (4 bytes per texel instead of 12); bit 31 of the width marks them
        assert scene.sky_rgbe.shape == (sh, sw, 4) and np.array_equal(decode_rgbe(scene.sky_rgbe).reshape(-1, 3), np.asarray(sky, np.float32).reshape(-1, 3))
        parts.append(struct.pack("<II", sw | 0x80000000, sh))
        put(scene.sky_rgbe.astype(np.uint8))
    else:
        parts.append(struct.pack("<II", sw, sh))
        put(np.asarray(sky, np.float32))
    put(scene.point_lights), put(scene.spot_lights), put(scene.dir_lights)
    raw = b"".join(parts)
    Path(path).parent.mkdir(parents=True, exist_ok=True)
    Path(path).write_bytes(BAKE_MAGIC + struct.pack("<Q", len(raw)) + zlib.compress(raw, 6))


def load_baked(path: Path) -> Scene:
    blob = Path(path).read_bytes()
    assert blob[:8] == BAKE_MAGIC
    (n,) = struct.unpack_from("<Q", blob, 8)
    raw = zlib.decompress(blob[16:])
    assert len(raw) == n
    off = 0

    def take(fmt):
        nonlocal off
        v = struct.unpack_from(fmt, raw, off)
        off += struct.calcsize(fmt)
        return v

    def get(dtype):
        nonlocal off
        (nb,) = take("<Q")
        a = np.frombuffer(raw, dtype=dtype, count=nb // np.dtype(dtype).itemsize, offset=off).copy()
        off += nb
        return a

    magic, nm, ni, nmat, ntex = take("<8sIIII")
    (ln,) = take("<I")
    name = raw[off:off + ln].decode()
    off += ln
    cam = take("<7f")
    s = Scene(name=name)
    s.camera_pos, s.camera_dir, s.fov = cam[0:3], cam[3:6], cam[6]
    for _ in range(nm):
        nv, nt, has_idx = take("<III")
        v = get(np.float32).reshape(nv, 4)
        t = get(R.TRIANGLE_DTYPE)
        idx = get(np.uint32).reshape(nt, 3) if has_idx else None
        s.meshes.append(SceneMesh(v, t, idx))
    for _ in range(ni):
        (mi,) = take("<I")
        s.instances.append((mi, get(np.float64).reshape(4, 4)))
    s.materials = get(R.MATERIAL_DTYPE)
    s.tex_ids = get(np.int32).reshape(-1, 11)
    for _ in range(ntex):
        ty, w, h = take("<III")
        data = get(np.uint32 if ty == R.TEX_UINT else np.float32)
        if ty == R.TEX_UINT and data.size == w * h:
            data = build_mips(data.reshape(h, w))
        s.textures.append({"type": ty, "width": w, "height": h, "data": data})
    sw, sh = take("<II")
    if sw & 0x80000000:
        sw &= 0x7FFFFFFF
        s.sky_rgbe = get(np.uint8).reshape(sh, sw, 4)
        s.sky = (decode_rgbe(s.sky_rgbe).reshape(-1, 3), sw, sh)
    else:
        s.sky = (get(np.float32).reshape(-1, 3), sw, sh)
    s.point_lights, s.spot_lights, s.dir_lights = get(R.POINT_LIGHT_DTYPE), get(R.SPOT_LIGHT_DTYPE), get(R.DIR_LIGHT_DTYPE)
    return s


BAKED_SPONZA = R.PKG_DIR / "data" / "_baked" / "sponza.rfwscene"


def sponza_or_standin() -> Scene:
    """config 2 scene: the reference's Sponza when tools/bake_sponza.py has produced it, else the atrium."""
    if BAKED_SPONZA.exists():
        return load_baked(BAKED_SPONZA)
    return atrium()


def sponza_instanced(copies: int = 38, seed: int = 1234) -> Scene:
    """BASELINE.json configs[2] stand-in (San Miguel is not in the reference repository): the
    config-2 scene instanced `copies` times on a jittered 3-D lattice with a rotation about Y per copy — 38 copies of
    Sponza's 262 k triangles are 10 M triangles behind one instance table, with the triangle-size distribution and the
    occlusion depth of the real asset.  Copy 0 is the config-2 scene itself (same light, same camera inside it)."""
    base = sponza_or_standin()
    if copies <= 1:
        return base
    rng = np.random.default_rng(seed)
    lo, hi = np.full(3, np.inf), np.full(3, -np.inf)
    for mi, M in base.instances:
        v = base.meshes[mi].vertices[:, :3].astype(np.float64)
        w = v @ np.asarray(M, np.float64)[:3, :3].T + np.asarray(M, np.float64)[:3, 3]
        lo, hi = np.minimum(lo, w.min(0)), np.maximum(hi, w.max(0))
    size = (hi - lo) * 1.05
    nx = int(np.ceil(copies ** (1 / 3) * 1.3))
    nz = int(np.ceil(np.sqrt(copies / nx * 1.5)))
    ny = int(np.ceil(copies / (nx * nz)))
    cells = [(x, y, z) for y in range(ny) for z in range(nz) for x in range(nx)]
    base_instances = list(base.instances)
    n_light = 1  # the light quad is the last instance of the baked scene and stays with copy 0 only
    for c in range(1, copies):
        x, y, z = cells[c]
        jitter = (rng.random(3) - 0.5) * 0.1 * size
        T = translate(*(np.array([x, y, z]) * size + jitter)) @ rotate_y(float(rng.integers(0, 4)) * 90.0 + float(rng.random() * 10 - 5))
        for mi, M in base_instances[: len(base_instances) - n_light]:
            base.instances.append((mi, T @ np.asarray(M, np.float64)))
    base.name = f"{base.name} x{copies} (instanced lattice, seed {seed})"
    return base


def add_config5_lights(scene: Scene, scene_scale: float = 0.2) -> Scene:
    """BASELINE.json configs[4]: add_point_light((-15, 10, -5) * s, (20, 20, 20)) and
    add_directional_light(normalize(-0.3, -1, 0.2), (3, 3, 3)) with energy = |radiance| (RFW/system/src/rfw/system.cpp:720-758)."""
    pl = np.zeros(1, R.POINT_LIGHT_DTYPE)
    pl["position"], pl["radiance"] = (-15 * scene_scale, 10 * scene_scale, -5 * scene_scale), (20, 20, 20)
    pl["energy"] = np.linalg.norm(pl["radiance"][0])
    dl = np.zeros(1, R.DIR_LIGHT_DTYPE)
    d = np.array([-0.3, -1.0, 0.2])
    dl["direction"], dl["radiance"] = d / np.linalg.norm(d), (3, 3, 3)
    dl["energy"] = np.linalg.norm(dl["radiance"][0])
    scene.point_lights, scene.dir_lights = pl, dl
    return scene
