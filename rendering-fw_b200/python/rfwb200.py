"""ctypes binding of the C ABI in include/rfwb200.h, used by tests/, bench.py and __graft_entry__.py.

`Library` is a (path, symbol prefix) pair, so the test infrastructure can drive a second library with the same call
shapes through the same wrapper class (the CPU checker is loaded by oracle/oracle_lib.py, never from this package).

`RenderContext` mirrors the method names of the reference's plugin interface
rfw::RenderContext (RFW/system/context/rfw/context/context.h:74-111); `Camera.get_view` restates
rfw::Camera::get_view (RFW/system/context/rfw/context/Camera.cpp:74-88,109-115).
"""
from __future__ import annotations

import ctypes as C
import math
import os
from pathlib import Path

import numpy as np

PKG_DIR = Path(__file__).resolve().parent.parent
REPO_DIR = PKG_DIR.parent
PRODUCT_LIB = Path(os.environ.get("RFWB200_LIB", PKG_DIR / "librfwb200.so"))  # override: tuning experiments only
BLUENOISE_BIN = PKG_DIR / "data" / "bluenoise_256spp.bin"

RESET, CONVERGE = 0, 1
TEX_FLOAT4, TEX_UINT = 0, 1

# ---- numpy dtypes of the wire formats (byte-exact, see include/rfwb200.h) -----------------------
TRIANGLE_DTYPE = np.dtype(
    [
        ("u", "<f4", 3), ("light_tri_idx", "<i4"),
        ("v", "<f4", 3), ("material", "<u4"),
        ("vN0", "<f4", 3), ("Nx", "<f4"),
        ("vN1", "<f4", 3), ("Ny", "<f4"),
        ("vN2", "<f4", 3), ("Nz", "<f4"),
        ("T", "<f4", 3), ("area", "<f4"),
        ("B", "<f4", 3), ("LOD", "<f4"),
        ("vertex0", "<f4", 3), ("dummy1", "<f4"),
        ("vertex1", "<f4", 3), ("dummy2", "<f4"),
        ("vertex2", "<f4", 3), ("dummy3", "<f4"),
    ]
)
MAP_DTYPE = np.dtype(
    [("width", "<i2"), ("height", "<i2"), ("uscale", "<f2"), ("vscale", "<f2"), ("uoffs", "<f2"), ("voffs", "<f2"), ("texaddr", "<u4")]
)
MATERIAL_DTYPE = np.dtype(
    [
        ("diffuse", "<f2", 3), ("transmittance", "<f2", 3), ("flags", "<u4"), ("parameters", "<u4", 4),
        ("tex0", MAP_DTYPE), ("tex1", MAP_DTYPE), ("tex2", MAP_DTYPE),
        ("nmap0", MAP_DTYPE), ("nmap1", MAP_DTYPE), ("nmap2", MAP_DTYPE),
        ("smap", MAP_DTYPE), ("rmap", MAP_DTYPE), ("cmap", MAP_DTYPE), ("amap", MAP_DTYPE),
    ]
)
AREA_LIGHT_DTYPE = np.dtype(
    [
        ("position", "<f4", 3), ("energy", "<f4"), ("normal", "<f4", 3), ("area", "<f4"),
        ("radiance", "<f4", 3), ("dummy0", "<i4"), ("vertex0", "<f4", 3), ("tri_idx", "<i4"),
        ("vertex1", "<f4", 3), ("inst_idx", "<i4"), ("vertex2", "<f4", 3), ("dummy1", "<i4"),
    ]
)
POINT_LIGHT_DTYPE = np.dtype([("position", "<f4", 3), ("energy", "<f4"), ("radiance", "<f4", 3), ("dummy", "<i4")])
SPOT_LIGHT_DTYPE = np.dtype(
    [("position", "<f4", 3), ("cos_inner", "<f4"), ("radiance", "<f4", 3), ("cos_outer", "<f4"), ("direction", "<f4", 3), ("energy", "<f4")]
)
DIR_LIGHT_DTYPE = np.dtype([("direction", "<f4", 3), ("energy", "<f4"), ("radiance", "<f4", 3), ("dummy", "<i4")])
HIT_DTYPE = np.dtype([("t", "<f4"), ("u", "<f4"), ("v", "<f4"), ("inst_id", "<i4"), ("prim_id", "<i4")])
assert TRIANGLE_DTYPE.itemsize == 160 and MATERIAL_DTYPE.itemsize == 192 and AREA_LIGHT_DTYPE.itemsize == 96
assert POINT_LIGHT_DTYPE.itemsize == 32 and SPOT_LIGHT_DTYPE.itemsize == 48 and DIR_LIGHT_DTYPE.itemsize == 32
assert HIT_DTYPE.itemsize == 20

# MatPropFlags bits, RFW/system/context/rfw/context/structs.h:67-83
MAT_HAS_DIFFUSE_MAP, MAT_HAS_NORMAL_MAP, MAT_SMOOTH_NORMALS, MAT_HAS_ALPHA = 1 << 2, 1 << 3, 1 << 11, 1 << 12


class TextureData(C.Structure):
    _fields_ = [("type", C.c_int32), ("width", C.c_uint32), ("height", C.c_uint32), ("texel_count", C.c_uint32),
                ("tex_addr", C.c_uint32), ("data", C.c_void_p)]


class Mesh(C.Structure):
    _fields_ = [("vertices", C.c_void_p), ("normals", C.c_void_p), ("tex_coords", C.c_void_p), ("triangles", C.c_void_p),
                ("indices", C.c_void_p), ("vertex_count", C.c_size_t), ("triangle_count", C.c_size_t)]


class LightCount(C.Structure):
    _fields_ = [("area", C.c_uint32), ("point", C.c_uint32), ("spot", C.c_uint32), ("directional", C.c_uint32)]


class CameraView(C.Structure):
    _fields_ = [("pos", C.c_float * 3), ("p1", C.c_float * 3), ("p2", C.c_float * 3), ("p3", C.c_float * 3),
                ("aperture", C.c_float), ("spread_angle", C.c_float)]


class RenderStats(C.Structure):
    _fields_ = [("primary_time", C.c_float), ("primary_count", C.c_uint32), ("secondary_time", C.c_float),
                ("secondary_count", C.c_uint32), ("deep_time", C.c_float), ("deep_count", C.c_uint32),
                ("shadow_time", C.c_float), ("shadow_count", C.c_uint32), ("shade_time", C.c_float),
                ("finalize_time", C.c_float), ("animation_time", C.c_float), ("render_time", C.c_float)]


class FrameCounters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("n_gen", "n_ext", "n_shade", "n_ext_out", "n_nee", "n_acc", "pixels", "samples")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}

    def algorithmic_bytes(self) -> int:
        """SURVEY.md §8(d) / BASELINE.md: compulsory bytes of the wavefront representation."""
        return (self.n_gen * 32 + self.n_ext * 48 + self.n_shade * 224 + self.n_ext_out * 48 + self.n_nee * 96 +
                self.n_acc * 32 + self.pixels * 32)


class GeometryStats(C.Structure):
    _fields_ = [("on_device", C.c_int32), ("was_refit", C.c_int32), ("device_ms", C.c_float), ("host_ms", C.c_float),
                ("refits", C.c_uint64), ("builds", C.c_uint64)]


BVH_NODE_DTYPE = np.dtype([("minx", "<f4", 4), ("maxx", "<f4", 4), ("miny", "<f4", 4), ("maxy", "<f4", 4), ("minz", "<f4", 4),
                           ("maxz", "<f4", 4), ("child", "<i4", 4), ("pad", "<i4", 4)])
TRI_REC_DTYPE = np.dtype([("p0", "<f4", 3), ("e1", "<f4", 3), ("e2", "<f4", 3), ("shade_idx", "<u4"), ("det_eps", "<f4"), ("pad0", "<u4")])
SHADE_TRI_DTYPE = np.dtype([("u", "<f4", 3), ("light_tri_idx", "<i4"), ("v", "<f4", 3), ("material", "<u4"),
                            ("n0", "<f4", 3), ("Nx", "<f4"), ("n1", "<f4", 3), ("Ny", "<f4"), ("n2", "<f4", 3), ("Nz", "<f4"),
                            ("area", "<f4"), ("lod", "<f4"), ("inst_id", "<u4"), ("prim_id", "<u4")])
assert BVH_NODE_DTYPE.itemsize == 128 and TRI_REC_DTYPE.itemsize == 48 and SHADE_TRI_DTYPE.itemsize == 96

assert C.sizeof(TextureData) == 32 and C.sizeof(Mesh) == 56 and C.sizeof(CameraView) == 56 and C.sizeof(RenderStats) == 48


class Rfwb200Error(RuntimeError):
    """The reference reports backend errors as std::runtime_error (CUDART/src/CheckCUDA.h:7-20)."""


def _f32(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if shape is not None:
        a = a.reshape(shape)
    return a


class Library:
    """One loaded shared library + its symbol prefix."""

    def __init__(self, path: Path, prefix: str):
        if not Path(path).exists():
            raise Rfwb200Error(f"{path} is missing — build it first (python -c 'import __graft_entry__ as g; g.build()'); "
                               "there is no fallback path")
        self.path = Path(path)
        self.prefix = prefix
        self.lib = C.CDLL(str(path))
        self.post_create = None  # optional callable(RenderContext) run after create (set by whoever loaded the library)

    def fn(self, name, restype=C.c_int, argtypes=None):
        f = getattr(self.lib, self.prefix + name)
        f.restype = restype
        if argtypes is not None:
            f.argtypes = argtypes
        return f

    def last_error(self) -> str:
        f = self.fn("last_error", C.c_char_p, [])
        return (f() or b"").decode()


def load_product() -> Library:
    return Library(PRODUCT_LIB, "rfwb200_")


class Camera:
    """rfw::Camera public fields (context/camera.h:27-37) + get_view (Camera.cpp:74-88)."""

    def __init__(self, position=(0, 0, 0), direction=(0, 0, 1), fov=40.0, width=1, height=1, focal_distance=5.0,
                 aperture=0.0001):
        self.position = np.asarray(position, np.float32)
        d = np.asarray(direction, np.float32)
        self.direction = d / np.float32(np.linalg.norm(d))
        self.FOV = np.float32(fov)
        self.focalDistance = np.float32(focal_distance)
        self.aperture = np.float32(aperture)
        self.pixelCount = (int(width), int(height))
        self.aspectRatio = np.float32(width) / np.float32(height)

    def get_view(self) -> CameraView:
        f32 = np.float32
        z = self.direction
        x = np.cross(z, np.array([0, 1, 0], f32)).astype(f32)
        x = x / f32(np.linalg.norm(x))
        y = np.cross(x, z).astype(f32)
        right, up, forward = x, y, z
        v = CameraView()
        spread = f32(self.FOV * f32(math.pi) / f32(180)) / f32(self.pixelCount[1])
        screen = f32(math.tan(float(self.FOV / f32(2.0) / f32(180.0 / math.pi))))
        center = self.position + self.focalDistance * forward
        p1 = center - screen * right * self.focalDistance * self.aspectRatio + screen * self.focalDistance * up
        p2 = center + screen * right * self.focalDistance * self.aspectRatio + screen * self.focalDistance * up
        p3 = center - screen * right * self.focalDistance * self.aspectRatio - screen * self.focalDistance * up
        for i in range(3):
            v.pos[i] = float(self.position[i])
            v.p1[i], v.p2[i], v.p3[i] = float(p1[i]), float(p2[i]), float(p3[i])
        v.aperture = float(self.aperture)
        v.spread_angle = float(spread)
        return v


class RenderContext:
    """Python mirror of rfw::RenderContext over the C ABI; method names follow context.h:78-110."""

    def __init__(self, library: Library, device: int = 0, devices=None):
        """devices=[0, 1, ...]: one context that owns several GPUs in this process (rfwb200_create_group) and shards every
        frame over them; otherwise `device` is the single CUDA ordinal."""
        self.L = library
        self._h = C.c_void_p()
        self._keep = []  # arrays borrowed by the last upload calls (API contract: caller keeps them alive during the call)
        if devices is not None:
            arr = (C.c_int * len(devices))(*[int(d) for d in devices])
            self._check(self.L.fn("create_group", C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)])(arr, len(devices), C.byref(self._h)))
        else:
            self._check(self.L.fn("create", C.c_int, [C.c_int, C.POINTER(C.c_void_p)])(device, C.byref(self._h)))
        self.width = self.height = 0
        if self.L.post_create is not None:
            self.L.post_create(self)

    # -- plumbing --
    def _check(self, rc: int):
        if rc != 0:
            raise Rfwb200Error(f"{self.L.prefix}* failed ({rc}): {self.L.last_error()}")

    def close(self):
        if self._h:
            self.L.fn("destroy", C.c_int, [C.c_void_p])(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- rfw::RenderContext --
    def init(self, width: int, height: int):
        self._check(self.L.fn("init", C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32])(self._h, width, height))
        self.width, self.height = width, height

    def set_shard(self, rank: int, world: int, tile_w: int = 32, tile_h: int = 8):
        self._check(self.L.fn("set_shard", C.c_int, [C.c_void_p] + [C.c_uint32] * 4)(self._h, rank, world, tile_w, tile_h))

    def set_stream(self, stream_handle: int):
        self._check(self.L.fn("set_stream", C.c_int, [C.c_void_p, C.c_void_p])(self._h, stream_handle))

    def set_sky(self, pixels_rgb: np.ndarray, width: int, height: int):
        px = _f32(pixels_rgb, (height * width, 3))
        self._check(self.L.fn("set_sky", C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t])(
            self._h, px.ctypes.data, width, height))

    def set_textures(self, textures):
        """textures: list of dicts {type, width, height, data(np array incl. mips for UINT)}"""
        n = len(textures)
        arr = (TextureData * max(n, 1))()
        keep = []
        for i, t in enumerate(textures):
            data = np.ascontiguousarray(t["data"], dtype=np.uint32 if t["type"] == TEX_UINT else np.float32)
            keep.append(data)
            arr[i].type = t["type"]
            arr[i].width, arr[i].height = t["width"], t["height"]
            arr[i].texel_count = data.size if t["type"] == TEX_UINT else data.size // 4
            arr[i].tex_addr = i
            arr[i].data = data.ctypes.data
        self._check(self.L.fn("set_textures", C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t])(self._h, C.addressof(arr), n))

    def set_materials(self, materials: np.ndarray, tex_ids: np.ndarray):
        m = np.ascontiguousarray(materials, dtype=MATERIAL_DTYPE)
        ids = np.ascontiguousarray(tex_ids, dtype=np.int32).reshape(len(m), 11)
        self._check(self.L.fn("set_materials", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t])(
            self._h, m.ctypes.data, ids.ctypes.data, len(m)))

    def set_mesh(self, index: int, vertices: np.ndarray, triangles: np.ndarray, indices: np.ndarray | None = None):
        v = _f32(vertices, (-1, 4))
        t = np.ascontiguousarray(triangles, dtype=TRIANGLE_DTYPE)
        m = Mesh()
        m.vertices, m.triangles = v.ctypes.data, t.ctypes.data
        m.normals = m.tex_coords = None
        idx = None
        if indices is not None:
            idx = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1, 3)
            m.indices = idx.ctypes.data
            assert len(idx) == len(t)
        else:
            m.indices = None
            assert len(v) == 3 * len(t)
        m.vertex_count, m.triangle_count = len(v), len(t)
        self._check(self.L.fn("set_mesh", C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p])(self._h, index, C.byref(m)))

    def set_instance(self, i: int, mesh_idx: int, transform: np.ndarray, normal_matrix: np.ndarray | None = None):
        """transform: 4x4 in the mathematical (row, col) convention; sent column-major like glm."""
        tm = np.asarray(transform, np.float64).reshape(4, 4)
        if normal_matrix is None:
            normal_matrix = np.linalg.inv(tm[:3, :3]).T  # mat3(transpose(inverse(transform))), system.cpp:347
        t = _f32(tm.T.reshape(-1))
        nm = _f32(np.asarray(normal_matrix, np.float64).reshape(3, 3).T.reshape(-1))
        self._check(self.L.fn("set_instance", C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p])(
            self._h, i, mesh_idx, t.ctypes.data, nm.ctypes.data))

    def set_lights(self, area=None, point=None, spot=None, directional=None):
        def prep(a, dt):
            a = np.zeros(0, dt) if a is None else np.ascontiguousarray(a, dtype=dt)
            return a, (a.ctypes.data if len(a) else None)
        a, ap = prep(area, AREA_LIGHT_DTYPE)
        p, pp = prep(point, POINT_LIGHT_DTYPE)
        s, sp = prep(spot, SPOT_LIGHT_DTYPE)
        d, dp = prep(directional, DIR_LIGHT_DTYPE)
        cnt = LightCount(len(a), len(p), len(s), len(d))
        self._check(self.L.fn("set_lights", C.c_int, [C.c_void_p, LightCount] + [C.c_void_p] * 4)(self._h, cnt, ap, pp, sp, dp))

    def update(self):
        self._check(self.L.fn("update", C.c_int, [C.c_void_p])(self._h))

    def set_setting(self, key: str, value):
        self._check(self.L.fn("set_setting", C.c_int, [C.c_void_p, C.c_char_p, C.c_char_p])(
            self._h, key.encode(), str(value).encode()))

    def get_settings(self) -> str:
        """key=value lines (RenderContext::get_settings, context.h:96) + `levels_in_use`: the form of the committed scene"""
        buf = C.create_string_buffer(4096)
        self._check(self.L.fn("get_settings", C.c_int, [C.c_void_p, C.c_char_p, C.c_size_t])(self._h, buf, len(buf)))
        return buf.value.decode()

    def render_frame(self, camera, status: int = RESET):
        view = camera.get_view() if hasattr(camera, "get_view") else camera
        self._check(self.L.fn("render_frame", C.c_int, [C.c_void_p, C.POINTER(CameraView), C.c_int])(
            self._h, C.byref(view), status))

    def local_pixel_count(self) -> int:
        return int(self.L.fn("local_pixel_count", C.c_size_t, [C.c_void_p])(self._h))

    def read_framebuffer(self, out: np.ndarray | None = None) -> np.ndarray:
        n = self.local_pixel_count()
        if out is None:
            out = np.empty((n, 4), np.float32)
        self._check(self.L.fn("read_framebuffer", C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t])(self._h, out.ctypes.data, n))
        return out

    def read_framebuffer_async(self, pinned_ptr: int, capacity_pixels: int):
        """rfwb200_read_framebuffer_async into pinned host memory at `pinned_ptr` (e.g. a torch pin_memory tensor's data_ptr)"""
        self._check(self.L.fn("read_framebuffer_async", C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t])(self._h, pinned_ptr, capacity_pixels))

    def read_wait(self):
        self._check(self.L.fn("read_wait", C.c_int, [C.c_void_p])(self._h))

    def read_aov(self, which: int) -> np.ndarray:
        """rfwb200_read_aov: 0 = albedo, 1 = normal (depth-0 feature planes, setting "aov")"""
        n = self.width * self.height
        out = np.empty((n, 4), np.float32)
        self._check(self.L.fn("read_aov", C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t])(self._h, which, out.ctypes.data, n))
        return out.reshape(self.height, self.width, 4)

    def set_aov_transform(self, m3: np.ndarray):
        m = _f32(np.asarray(m3, np.float64).reshape(3, 3).T.reshape(-1))
        self._check(self.L.fn("set_aov_transform", C.c_int, [C.c_void_p, C.c_void_p])(self._h, m.ctypes.data))

    def read_image(self) -> np.ndarray:
        return self.read_framebuffer().reshape(self.height, self.width, 4)

    def read_display(self, contrast: float = 0.0, brightness: float = 0.0, out: np.ndarray | None = None) -> np.ndarray:
        """rfwb200_read_display: the tone-map pass of system::render_frame(toneMap=true) -> (local pixels, 4) uint8"""
        n = self.local_pixel_count()
        if out is None:
            out = np.empty((n, 4), np.uint8)
        assert out.dtype == np.uint8 and out.size >= n * 4 and out.flags.c_contiguous
        self._check(self.L.fn("read_display", C.c_int, [C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_size_t])(
            self._h, contrast, brightness, out.ctypes.data, n))
        return out

    def set_probe_index(self, x: int, y: int):
        self._check(self.L.fn("set_probe_index", C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32])(self._h, x, y))

    def get_probe_results(self):
        i, p, d = C.c_uint32(), C.c_uint32(), C.c_float()
        self._check(self.L.fn("get_probe_results", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p])(
            self._h, C.byref(i), C.byref(p), C.byref(d)))
        return int(i.value), int(p.value), float(d.value)

    def get_frame_counters(self) -> FrameCounters:
        fc = FrameCounters()
        self._check(self.L.fn("get_frame_counters", C.c_int, [C.c_void_p, C.c_void_p])(self._h, C.byref(fc)))
        return fc

    # -- product-only --
    def synchronize(self):
        self._check(self.L.fn("synchronize", C.c_int, [C.c_void_p])(self._h))

    def device_framebuffer(self) -> int:
        return int(self.L.fn("device_framebuffer", C.c_void_p, [C.c_void_p])(self._h) or 0)

    def shard_stride(self) -> int:
        return int(self.L.fn("shard_stride", C.c_size_t, [C.c_void_p])(self._h))

    # -- display image of a sharded frame, one rank per process (rfwb200_display_*) --
    def display_create(self) -> int:
        ptr = C.c_void_p()
        self._check(self.L.fn("display_create", C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)])(self._h, C.byref(ptr)))
        return int(ptr.value or 0)

    def display_export(self) -> bytes:
        buf = (C.c_ubyte * 64)()
        self._check(self.L.fn("display_export", C.c_int, [C.c_void_p, C.c_void_p])(self._h, buf))
        return bytes(buf)

    def display_import(self, handle: bytes):
        assert len(handle) == 64
        buf = (C.c_ubyte * 64)(*handle)
        self._check(self.L.fn("display_import", C.c_int, [C.c_void_p, C.c_void_p])(self._h, buf))

    def display_attach(self, display_rank: "RenderContext"):
        self._check(self.L.fn("display_attach", C.c_int, [C.c_void_p, C.c_void_p])(self._h, display_rank._h))

    def display_wait(self):
        self._check(self.L.fn("display_wait", C.c_int, [C.c_void_p])(self._h))

    def display_image(self) -> int:
        return int(self.L.fn("display_image", C.c_void_p, [C.c_void_p])(self._h) or 0)

    def debug_read_plane(self, which: int, n: int) -> np.ndarray:
        """wavefront plane of the last frame (rfwb200_debug_read_plane): 6 = hit records (bits(w0_16 | w1_16 << 16), bits(shading
        record), bits(prim or -1), t) per work item"""
        out = np.empty((n, 4), np.float32)
        self._check(self.L.fn("debug_read_plane", C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t])(self._h, which, out.ctypes.data, n))
        return out

    def read_device(self, ptr: int, n_pixels: int) -> np.ndarray:
        """blocking copy of n_pixels float4 at a device pointer of this context's device (tests)"""
        out = np.empty((n_pixels, 4), np.float32)
        self._check(self.L.fn("debug_read_device", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t])(self._h, ptr, out.ctypes.data, out.nbytes))
        return out

    def assemble_shards(self, gathered_ptr: int, image_ptr: int):
        self._check(self.L.fn("assemble_shards", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p])(self._h, gathered_ptr, image_ptr))

    def get_stats(self) -> RenderStats:
        s = RenderStats()
        self._check(self.L.fn("get_stats", C.c_int, [C.c_void_p, C.c_void_p])(self._h, C.byref(s)))
        return s

    def launch_count(self) -> int:
        return int(self.L.fn("launch_count", C.c_uint64, [C.c_void_p])(self._h))

    def get_bvh_info(self):
        n, t, s, ms = C.c_uint64(), C.c_uint64(), C.c_float(), C.c_float()
        self._check(self.L.fn("get_bvh_info", C.c_int, [C.c_void_p] + [C.c_void_p] * 4)(
            self._h, C.byref(n), C.byref(t), C.byref(s), C.byref(ms)))
        return {"nodes": int(n.value), "triangles": int(t.value), "sah_cost": float(s.value), "build_ms": float(ms.value)}

    # -- device geometry (product-only extensions: GPU skinning + refit) --
    def set_mesh_skin(self, mesh_index: int, base_vertices, base_normals, joints, weights):
        """bind pose of a skinned mesh: vec4 vertices, vec3/vec4 normals, uvec4 joints, vec4 weights per vertex
        (rfw::geometry::gltf::SceneMesh::baseVertices/baseNormals/joints/weights)."""
        v = _f32(base_vertices, (-1, 4))
        n = np.asarray(base_normals, np.float32)
        n = n.reshape(len(v), -1)
        n4 = np.zeros((len(v), 4), np.float32)
        n4[:, : min(n.shape[1], 3)] = n[:, :3]
        j = np.ascontiguousarray(joints, dtype=np.uint32).reshape(len(v), 4)
        w = _f32(weights, (len(v), 4))
        self._check(self.L.fn("set_mesh_skin", C.c_int, [C.c_void_p, C.c_size_t] + [C.c_void_p] * 4 + [C.c_size_t])(
            self._h, mesh_index, v.ctypes.data, n4.ctypes.data, j.ctypes.data, w.ctypes.data, len(v)))

    def set_mesh_pose(self, mesh_index: int, joint_matrices):
        """joint_matrices: (n_joints, 4, 4) in the mathematical (row, col) convention; sent column-major like glm."""
        m = np.asarray(joint_matrices, np.float32).reshape(-1, 4, 4)
        cm = np.ascontiguousarray(np.transpose(m, (0, 2, 1)))
        self._check(self.L.fn("set_mesh_pose", C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t])(
            self._h, mesh_index, cm.ctypes.data, len(cm)))

    def set_mesh_morph_targets(self, mesh_index: int, pose_positions, pose_normals):
        """poses: (n_targets + 1, nv, 3|4); pose 0 is the base (SceneMesh::poses, gltf/mesh.h)."""
        def as4(a):
            a = np.asarray(a, np.float32)
            out = np.zeros(a.shape[:2] + (4,), np.float32)
            out[..., :3] = a[..., :3]
            return np.ascontiguousarray(out)
        p, n = as4(pose_positions), as4(pose_normals)
        self._check(self.L.fn("set_mesh_morph_targets", C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t])(
            self._h, mesh_index, p.ctypes.data, n.ctypes.data, p.shape[0] - 1, p.shape[1]))

    def set_mesh_morph_weights(self, mesh_index: int, weights):
        w = _f32(weights, (-1,))
        self._check(self.L.fn("set_mesh_morph_weights", C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t])(
            self._h, mesh_index, w.ctypes.data, len(w)))

    def get_geometry_stats(self) -> GeometryStats:
        g = GeometryStats()
        self._check(self.L.fn("get_geometry_stats", C.c_int, [C.c_void_p, C.c_void_p])(self._h, C.byref(g)))
        return g

    def debug_read_scene(self, which: str) -> np.ndarray:
        """'nodes' | 'tris' | 'shade' of the committed device scene, as structured arrays."""
        code, dt = {"nodes": (0, BVH_NODE_DTYPE), "tris": (1, TRI_REC_DTYPE), "shade": (2, SHADE_TRI_DTYPE)}[which]
        f = self.L.fn("debug_read_scene", C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p])
        nbytes = C.c_size_t()
        self._check(f(self._h, code, None, 0, C.byref(nbytes)))
        out = np.zeros(nbytes.value // dt.itemsize, dt)
        if nbytes.value:
            self._check(f(self._h, code, out.ctypes.data, out.nbytes, C.byref(nbytes)))
        return out

    # -- stage-level --
    def trace_closest(self, origins: np.ndarray, directions: np.ndarray, t_min: float = 1e-5) -> np.ndarray:
        o, d = _f32(origins, (-1, 4)), _f32(directions, (-1, 4))
        hits = np.zeros(len(o), HIT_DTYPE)
        self._check(self.L.fn("trace_closest", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_float, C.c_void_p])(
            self._h, o.ctypes.data, d.ctypes.data, len(o), t_min, hits.ctypes.data))
        return hits

    def trace_occluded(self, origins, directions, t_max, t_min: float = 1e-5) -> np.ndarray:
        o, d, tm = _f32(origins, (-1, 4)), _f32(directions, (-1, 4)), _f32(t_max, (-1,))
        occ = np.zeros(len(o), np.uint8)
        self._check(self.L.fn("trace_occluded", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_float, C.c_void_p])(
            self._h, o.ctypes.data, d.ctypes.data, tm.ctypes.data, len(o), t_min, occ.ctypes.data))
        return occ

    def generate_primary(self, camera, sample_index: int = 0):
        view = camera.get_view() if hasattr(camera, "get_view") else camera
        n = self.width * self.height
        o, d = np.zeros((n, 4), np.float32), np.zeros((n, 4), np.float32)
        self._check(self.L.fn("generate_primary", C.c_int, [C.c_void_p, C.POINTER(CameraView), C.c_uint32, C.c_void_p, C.c_void_p, C.c_size_t])(
            self._h, C.byref(view), sample_index, o.ctypes.data, d.ctypes.data, n))
        return o, d

    # -- entry points librfwb200.so does not export (a checker library may) --
    def intersect_prim(self, origin, direction, inst: int, prim: int, t_min: float = 1e-5) -> float:
        o, d = _f32(origin, (-1,)), _f32(direction, (-1,))
        t = C.c_float()
        self._check(self.L.fn("intersect_prim", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p])(
            self._h, o.ctypes.data, d.ctypes.data, inst, prim, t_min, C.byref(t)))
        return float(t.value)

    def last_render_ms(self) -> float:
        return float(self.L.fn("last_render_ms", C.c_float, [C.c_void_p])(self._h))


# ---- screen-tile sharding (host mirror of local_to_pixel in csrc/kernels.cu) -------
def shard_pixel_map(width: int, height: int, rank: int, world: int, tile_w: int = 32, tile_h: int = 8) -> np.ndarray:
    """global pixel id (y*width+x) of every local work index of `rank`, -1 for padded (dead) slots."""
    tiles_x, tiles_y = (width + tile_w - 1) // tile_w, (height + tile_h - 1) // tile_h
    total = tiles_x * tiles_y
    local_tiles = (total - rank + world - 1) // world if total > rank else 0
    tp = tile_w * tile_h
    j = np.arange(local_tiles * tp, dtype=np.int64)
    lt, w = j // tp, j % tp
    gt = lt * world + rank
    ty, tx = gt // tiles_x, gt % tiles_x
    blk, lane = w >> 5, w & 31
    bpr = tile_w >> 3
    by, bx = blk // bpr, blk % bpr
    x = tx * tile_w + bx * 8 + ((lane & 1) | ((lane >> 1) & 2) | ((lane >> 2) & 4))  # Z curve inside the 8x4 block (kernels.cu local_to_pixel)
    y = ty * tile_h + by * 4 + (((lane >> 1) & 1) | ((lane >> 2) & 2))
    ok = (x < width) & (y < height) & (ty < tiles_y)
    return np.where(ok, y * width + x, -1)


def shard_stride(width: int, height: int, world: int, tile_w: int = 32, tile_h: int = 8) -> int:
    tiles_x, tiles_y = (width + tile_w - 1) // tile_w, (height + tile_h - 1) // tile_h
    return ((tiles_x * tiles_y + world - 1) // world) * tile_w * tile_h


def assemble_shards_host(shards, width: int, height: int, tile_w: int = 32, tile_h: int = 8) -> np.ndarray:
    """shards[r]: (>= local pixels, 4) float32 tile-major framebuffer of rank r -> (height, width, 4) image."""
    world = len(shards)
    img = np.zeros((height * width, 4), np.float32)
    for r, sh in enumerate(shards):
        m = shard_pixel_map(width, height, r, world, tile_w, tile_h)
        ok = m >= 0
        img[m[ok]] = np.asarray(sh)[: len(m)][ok]
    return img.reshape(height, width, 4)
