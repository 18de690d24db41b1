#!/usr/bin/env python3
"""Generate the committed golden fixtures from the strict CPU oracle (oracle/librfworacle.so).

The reference holds no golden vectors for this path, so these fixtures pin OUR oracle:
they freeze its outputs on small seeded inputs so that (a) a change to the oracle is a visible diff and
(b) the GPU tests can also be checked against committed numbers.  Integer items (hashes, RNG streams,
blue-noise bytes, packed normals, tile maps) are exact; float images are stored as float32.
Run:  python tests/golden/make_golden.py     (rewrites tests/golden/*.npz)
"""
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(REPO / "rendering-fw_b200" / "python"))
sys.path.insert(0, str(REPO))
import ctypes as C  # noqa: E402

import rfwb200 as R  # noqa: E402
from oracle.oracle_lib import load_oracle  # noqa: E402
import scenes as S  # noqa: E402

OUT = Path(__file__).resolve().parent


def scalar_kats(lib):
    wang = lib.fn("wang_hash", C.c_uint32, [C.c_uint32])
    xor128 = lib.fn("xor128", C.c_uint32, [C.c_uint32, C.c_uint32])
    pack = lib.fn("pack_normal", C.c_uint32, [C.c_void_p])
    unpack = lib.fn("unpack_normal", None, [C.c_uint32, C.c_void_p])
    rint = lib.fn("random_int", C.c_uint32, [C.c_void_p])
    seeds = np.array([0, 1, 2, 61, 12345, 0xDEADBEEF, 0xFFFFFFFF, 16789, 1791, 720898027], np.uint32)
    out = {"wang_in": seeds, "wang_out": np.array([wang(int(s)) for s in seeds], np.uint32)}
    out["xor128_default_first8"] = np.array([xor128(123456789, n) for n in range(1, 9)], np.uint32)
    st = C.c_uint32(0x9E3779B9)
    out["xorshift_stream"] = np.array([rint(C.byref(st)) for _ in range(8)], np.uint32)
    rng = np.random.default_rng(1)
    n = rng.normal(size=(32, 3)).astype(np.float32)
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    n[:, 2] = np.abs(n[:, 2])
    packed = np.array([pack(n[i].ctypes.data) for i in range(len(n))], np.uint32)
    un = np.zeros((len(n), 3), np.float32)
    for i in range(len(n)):
        unpack(int(packed[i]), un[i].ctypes.data)
    out.update(normals=n, packed=packed, unpacked=un)
    return out


def main():
    lib = load_oracle()
    np.savez_compressed(OUT / "scalar_kats.npz", **scalar_kats(lib))

    W, H = 64, 48
    items = {}
    for name, fn in (("cornell", lambda: S.cornell_box(unit_scale=True)), ("soup", S.feature_soup)):
        sc = fn()
        ctx = R.RenderContext(lib)
        S.upload(ctx, sc, W, H)
        cam = sc.camera(W, H)
        bn = lib.fn("blue_noise", C.c_float, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int])
        if name == "cornell":
            items["blue_noise_samples"] = np.array([[bn(ctx._h, x, y, s, d) for d in range(6)] for x, y, s in
                                                    ((0, 0, 0), (5, 9, 1), (127, 127, 255), (130, 3, 256), (64, 32, 17))], np.float32)
        o, d = ctx.generate_primary(cam, 0)
        hits = ctx.trace_closest(o, d)
        items[f"{name}_origins"], items[f"{name}_dirs"], items[f"{name}_hits"] = o, d, hits
        for mode, depth in (("embree", 2), ("pt", 0), ("pt", 1), ("pt", 2)):
            ctx.set_setting("mode", mode)
            ctx.set_setting("max_path_length", depth)
            ctx.set_setting("spp", 1)
            ctx.render_frame(cam, R.RESET)
            items[f"{name}_{mode}_d{depth}"] = ctx.read_image().copy()
            items[f"{name}_{mode}_d{depth}_counters"] = np.array(list(ctx.get_frame_counters().as_dict().values()), np.uint64)
    np.savez_compressed(OUT / "oracle_frames_64x48.npz", **items)
    for f in sorted(OUT.glob("*.npz")):
        print(f.name, f.stat().st_size)


if __name__ == "__main__":
    main()
