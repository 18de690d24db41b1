"""CPU restatement (numpy, fp32) of the reference's glTF skinning — TEST INFRASTRUCTURE ONLY (never imported by the
product path; only tests/, __graft_entry__.smoke() and bench.py's cpu legs may use oracle/).

Follows RFW/system/src/rfw/geometry/gltf/mesh.cpp:
  * set_pose(skin)      :18-48   S = sum_k w_k * J[j_k];  v' = S * v;  n' = normalize((n^T * S^-1).xyz)
  * set_pose(weights)   :126-148 morph targets: v = pose0 + sum_j w_j pose_j (normals likewise, not renormalised)
  * update_triangles()  :428-449 vertex0..2 / vN0..2 from the indexed vertices, N = normalize(cross(v1-v0, v2-v0))
  * joint matrices      node.cpp:97-104  J[j] = inverse(meshNode.combined) * jointNode.combined * inverseBind[j]
Parity: the skinning step is pinned on the reference's own SIMD math (RFW/system/math/src/rfw/math.h compiled from the reference
tree around the loop body of mesh.cpp:30-45, oracle/_ref/librfwref_skin.so; vectors tests/golden/ref_skin_vectors.npz —
vertices bit for bit, normals within 1 ulp); morph targets and update_triangles are restated from source, unpinned (the
reference has no tests or vectors for them).  The GPU kernels (rendering-fw_b200/csrc/geometry.cu) are compared
against this file within 2e-5 relative and against the same reference vectors.
"""
from __future__ import annotations

import numpy as np


def skin_matrices(joints: np.ndarray, weights: np.ndarray, joint_matrices: np.ndarray) -> np.ndarray:
    """(nv,4) uint, (nv,4) f32, (nj,4,4) f32 (row, col) -> (nv,4,4) f32, accumulated in the reference's order."""
    J = np.asarray(joint_matrices, np.float32)
    w = np.asarray(weights, np.float32)
    j = np.asarray(joints, np.int64)
    S = J[j[:, 0]] * w[:, 0, None, None]
    S = S + J[j[:, 1]] * w[:, 1, None, None]
    S = S + J[j[:, 2]] * w[:, 2, None, None]
    S = S + J[j[:, 3]] * w[:, 3, None, None]
    return S.astype(np.float32)


def set_pose(base_vertices, base_normals, joints, weights, joint_matrices):
    """-> (vertices (nv,4) f32, normals (nv,3) f32)   [mesh.cpp:18-48]"""
    S = skin_matrices(joints, weights, joint_matrices)
    v = np.asarray(base_vertices, np.float32).reshape(-1, 4)
    n = np.asarray(base_normals, np.float32).reshape(len(v), -1)[:, :3]
    out_v = np.einsum("nij,nj->ni", S, v).astype(np.float32)
    Sinv = np.linalg.inv(S.astype(np.float64))
    n4 = np.concatenate([n, np.zeros((len(n), 1), np.float32)], 1).astype(np.float64)
    r = np.einsum("nj,nji->ni", n4, Sinv)[:, :3]  # row vector times matrix
    # the reference divides the whole vec4 by ITS length (w included); with an affine S and n.w = 0 the w term is 0
    r = r / np.linalg.norm(np.einsum("nj,nji->ni", n4, Sinv), axis=1, keepdims=True)
    return out_v, r.astype(np.float32)


def set_pose_morph(pose_positions, pose_normals, weights):
    """mesh.cpp:126-148: poses (n_targets + 1, nv, 3); -> (vertices (nv,4) with w = 1, normals (nv,3), NOT renormalised)"""
    P, N = np.asarray(pose_positions, np.float32)[..., :3], np.asarray(pose_normals, np.float32)[..., :3]
    w = np.asarray(weights, np.float32)
    assert len(w) == len(P) - 1
    v, n = P[0].copy(), N[0].copy()
    for j in range(len(w)):
        v = v + w[j] * P[j + 1]
        n = n + w[j] * N[j + 1]
    return np.concatenate([v, np.ones((len(v), 1), np.float32)], 1).astype(np.float32), n.astype(np.float32)


def update_triangles(triangles, vertices, normals, indices):
    """mesh.cpp:428-449 (indexed): returns a copy of the 160-B records with vertex0..2, vN0..2, N refreshed."""
    t = triangles.copy()
    idx = np.asarray(indices, np.int64).reshape(-1, 3)
    p = np.asarray(vertices, np.float32)[:, :3][idx]
    n = np.asarray(normals, np.float32)[idx]
    t["vertex0"], t["vertex1"], t["vertex2"] = p[:, 0], p[:, 1], p[:, 2]
    t["vN0"], t["vN1"], t["vN2"] = n[:, 0], n[:, 1], n[:, 2]
    cr = np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0]).astype(np.float32)
    N = cr / np.linalg.norm(cr, axis=1, keepdims=True).astype(np.float32)
    t["Nx"], t["Ny"], t["Nz"] = N[:, 0], N[:, 1], N[:, 2]
    return t
