#!/usr/bin/env python3
"""Run seeded box sets through the REFERENCE's own in-tree BVH node code (oracle/_ref/librfwref_bvh.so = RFW/system/bvh
{aabb,bvh_node,mbvh_node}.{h,cpp} compiled from /root/reference: BVHNode::subdivide<9, 32, 3>, MBVHNode::merge_nodes) and commit the
trees as tests/golden/ref_bvh_vectors.npz.  tests/test_ref_pin_bvh.py checks the oracle's builder against them everywhere and against
the live library where it exists.  Runs only in the build container (the reference does not travel)."""
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(REPO / "rendering-fw_b200" / "python"))
sys.path.insert(0, str(REPO / "tests"))
from ref_pin_bvh_common import box_sets, ref_build  # noqa: E402

OUT = Path(__file__).resolve().parent / "ref_bvh_vectors.npz"


def main():
    out = {}
    for name, boxes in box_sets().items():
        nodes, prims, mnodes = ref_build(boxes)
        out[f"{name}/boxes"], out[f"{name}/nodes"], out[f"{name}/prims"], out[f"{name}/mnodes"] = boxes, nodes.view(np.uint32).reshape(-1, 8), prims, mnodes
        print(name, len(boxes), "boxes ->", len(nodes), "nodes,", len(mnodes), "4-wide nodes")
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, OUT.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
