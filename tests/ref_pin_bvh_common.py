"""ctypes wrappers around oracle/_ref/librfwref_bvh.so (the reference's in-tree BVH node code compiled in place,
oracle/ref_build/ref_bvh_shim.cpp) and the matching oracle hook, plus the seeded box sets both are run on."""
import ctypes as C

import numpy as np

import rfwb200 as R

REF_BVH_LIB = R.REPO_DIR / "oracle" / "_ref" / "librfwref_bvh.so"
NODE_DTYPE = np.dtype([("bmin", "<f4", 3), ("bmax", "<f4", 3), ("left_first", "<i4"), ("count", "<i4")])  # bvh_node.h:23-28
assert NODE_DTYPE.itemsize == 32


def ref_build(aabbs):
    """reference: BVHNode::subdivide<9, 32, 3> + MBVHNode::merge_nodes -> (BVH2 nodes, primitive order, 4-wide nodes as 32 words each)"""
    lib = C.CDLL(str(REF_BVH_LIB))
    a = np.ascontiguousarray(aabbs, np.float32).reshape(-1, 6)
    n, cap = len(a), max(2 * len(a), 2)
    nodes, prims, mnodes = np.zeros(cap, NODE_DTYPE), np.zeros(n, np.uint32), np.zeros((cap, 32), np.uint32)
    nn, nm = C.c_int(), C.c_int()
    rc = lib.rfwref_bvh_build(C.c_void_p(a.ctypes.data), n, C.c_void_p(nodes.ctypes.data), C.c_void_p(prims.ctypes.data), C.byref(nn),
                              C.c_void_p(mnodes.ctypes.data), C.byref(nm))
    assert rc in (0, 1)
    return nodes[: nn.value].copy(), prims, mnodes[: nm.value].copy()


def oracle_build(oracle_lib, aabbs):
    a = np.ascontiguousarray(aabbs, np.float32).reshape(-1, 6)
    n, cap = len(a), max(2 * len(a), 2)
    nodes, prims, mnodes = np.zeros(cap, NODE_DTYPE), np.zeros(n, np.uint32), np.zeros((cap, 32), np.uint32)
    nn, nm = C.c_size_t(), C.c_size_t()
    f = oracle_lib.fn("build_bvh_from_aabbs", C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p])
    assert f(a.ctypes.data, n, nodes.ctypes.data, cap, prims.ctypes.data, mnodes.ctypes.data, cap, C.byref(nn), C.byref(nm)) == 0, oracle_lib.last_error()
    return nodes[: nn.value].copy(), prims, mnodes[: nm.value].copy()


def triangle_boxes(scene):
    """world-space boxes of every triangle of every instance of a scenes.Scene"""
    out = []
    for mesh_idx, M in scene.instances:
        m = scene.meshes[mesh_idx]
        v = np.asarray(m.vertices, np.float64)[:, :3]
        idx = np.asarray(m.indices, np.int64) if m.indices is not None else np.arange(len(v)).reshape(-1, 3)
        p = v[idx] @ np.asarray(M, np.float64)[:3, :3].T + np.asarray(M, np.float64)[:3, 3]
        out.append(np.concatenate([p.min(1), p.max(1)], 1))
    return np.concatenate(out).astype(np.float32)


def box_sets():
    """name -> (n, 6) float32 boxes: random, clustered, a regular grid (ties everywhere), duplicates, and two scenes' triangles"""
    import scenes as S

    rng = np.random.default_rng(2024)
    sets = {}
    for n in (1, 2, 3, 4, 17, 200, 2000):
        c = rng.uniform(-10, 10, (n, 3)).astype(np.float32)
        e = rng.uniform(0.01, 1.0, (n, 3)).astype(np.float32)
        sets[f"random{n}"] = np.concatenate([c - e, c + e], 1)
    c = np.concatenate([rng.normal(m, 0.3, (300, 3)) for m in ((-5, 0, 0), (5, 1, 0), (0, 0, 9))]).astype(np.float32)
    sets["clusters"] = np.concatenate([c - np.float32(0.05), c + np.float32(0.05)], 1)
    g = np.stack(np.meshgrid(np.arange(12), np.arange(12), np.arange(3), indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    sets["grid"] = np.concatenate([g, g + np.float32(1.0)], 1)  # equal costs on many planes: the first plane found must win on both sides
    d = np.repeat(sets["random17"], 9, axis=0)
    sets["duplicates"] = d  # nine copies of each box: planes that move nothing
    flat = sets["random200"].copy()
    flat[:, [1, 4]] = 0.0
    sets["flat"] = flat  # zero extent along y: zero-area candidates
    sets["cornell"] = triangle_boxes(S.cornell_box(unit_scale=True))
    sets["soup"] = triangle_boxes(S.feature_soup())
    return sets
