"""The oracle's BVH builder pinned on the reference's own in-tree builder code.

The reference's tree classes hand the build to the un-vendored Rust crate rtbvh (bvh_tree.cpp:48-102), which is absent here; the
builder the reference tree itself still holds is the node-level code of RFW/system/bvh — BVHNode::subdivide / partition (binned
SAH, bvh_node.h:56-81,136-233) and MBVHNode::merge_nodes / merge_node (4-wide collapse, mbvh_node.cpp:194-374) — and that is what
oracle/rfw_oracle.cpp restates.  oracle/ref_build compiles those headers and sources from where they lie into
oracle/_ref/librfwref_bvh.so; its trees over seeded box sets are committed (tests/golden/ref_bvh_vectors.npz, generator
make_ref_bvh_golden.py).  The oracle must produce the SAME trees: primitive order, every BVH2 node (bounds bit for bit, child /
first-primitive index, count) and every 4-wide node (all 128 bytes), on random, clustered, gridded (ties on every plane),
duplicated and flat boxes and on two scenes' triangles."""
from pathlib import Path

import numpy as np
import pytest

from ref_pin_bvh_common import NODE_DTYPE, REF_BVH_LIB, box_sets, oracle_build, ref_build

GOLD = Path(__file__).resolve().parent / "golden" / "ref_bvh_vectors.npz"


def same_tree(got, want, name):
    nodes, prims, mnodes = got
    wn, wp, wm = want
    assert np.array_equal(prims, wp), f"{name}: primitive order"
    assert len(nodes) == len(wn), f"{name}: {len(nodes)} nodes, reference {len(wn)}"
    # node 1 is never used by either side (children are allocated in pairs from index 2): its content is whatever the pool held
    used = np.ones(len(nodes), bool)
    if len(nodes) > 1:
        used[1] = False
    for k in ("bmin", "bmax"):
        assert np.array_equal(nodes[k][used].view(np.uint32), wn[k][used].view(np.uint32)), f"{name}: {k}"
    inner = (wn["count"] < 0) & used
    leaf = (wn["count"] >= 0) & used
    assert np.array_equal(nodes["count"][used], wn["count"][used]), f"{name}: counts"
    assert np.array_equal(nodes["left_first"][inner], wn["left_first"][inner]) and np.array_equal(nodes["left_first"][leaf], wn["left_first"][leaf]), f"{name}: child / first-primitive indices"
    if len(wm):  # (a root that stayed a leaf has no 4-wide tree in the reference: merge_nodes refuses leaves)
        assert np.array_equal(mnodes, wm), f"{name}: 4-wide nodes differ in {(mnodes != wm).any(1).sum()} of {len(wm)}"


def test_oracle_builder_equals_the_committed_reference_trees(oracle_lib):
    gold = np.load(GOLD)
    names = sorted({k.split("/")[0] for k in gold.files})
    assert len(names) >= 12
    for name in names:
        want = (gold[f"{name}/nodes"].view(NODE_DTYPE).reshape(-1), gold[f"{name}/prims"], gold[f"{name}/mnodes"])
        same_tree(oracle_build(oracle_lib, gold[f"{name}/boxes"]), want, name)


@pytest.mark.skipif(not REF_BVH_LIB.exists(), reason="oracle/_ref/librfwref_bvh.so is only built where /root/reference exists")
def test_committed_trees_are_what_the_live_reference_code_builds(oracle_lib):
    gold = np.load(GOLD)
    sets = box_sets()
    for name, boxes in sets.items():
        assert np.array_equal(boxes, gold[f"{name}/boxes"]), name
        nodes, prims, mnodes = ref_build(boxes)
        assert np.array_equal(nodes.view(np.uint32).reshape(-1, 8), gold[f"{name}/nodes"]) and np.array_equal(prims, gold[f"{name}/prims"]) and np.array_equal(mnodes, gold[f"{name}/mnodes"]), name
    # and a larger live case than the committed ones
    rng = np.random.default_rng(9)
    c = rng.uniform(-50, 50, (30_000, 3)).astype(np.float32)
    e = rng.uniform(0.01, 2.0, (30_000, 3)).astype(np.float32)
    boxes = np.concatenate([c - e, c + e], 1)
    same_tree(oracle_build(oracle_lib, boxes), ref_build(boxes), "random30000")
