"""CPU oracle (TEST INFRASTRUCTURE ONLY -- never imported by the product package) for SURVEY 8f ranks 1-2:
closest-point correspondences and deformation-graph maintenance, restated in numpy from the reference.

Pinned by tests/test_oracle_vs_reference.py (the unmodified reference executed live, marching cubes replaced by a
function returning the supplied live vertices -- scikit-image is not installed) and by the `corr_*` / `graph_*`
arrays of tests/golden/reference_vectors.npz.
"""
import numpy as np
from scipy.spatial import cKDTree

from . import dq as odq


def knn_points(points, queries, k):
    """KDTree(points).query(q, k) (core/fusion.py:255,264): ids ascending by float64 Euclidean distance.
    Returns (idx [m,k], d2 [m,k]) from brute force with a stable sort (ties: lower id first)."""
    p = np.asarray(points, dtype=np.float64)
    q = np.asarray(queries, dtype=np.float64)
    idx = np.empty((len(q), k), dtype=np.int64)
    d2o = np.empty((len(q), k))
    for s in range(0, len(q), 2048):
        d = q[s:s + 2048, None, :] - p[None, :, :]
        d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
        o = np.argsort(d2, axis=1, kind="stable")[:, :k]
        idx[s:s + 2048] = o
        d2o[s:s + 2048] = np.take_along_axis(d2, o, 1)
    return idx, d2o


def knn_tie(points, queries, k, rel=1e-12):
    """True where the k-th / (k+1)-th neighbour (or any adjacent pair among the k) is tied to rounding: the KD-tree's
    answer is then implementation-defined."""
    kk = min(k + 1, len(points))
    _, d2 = knn_points(points, queries, kk)
    gap = np.diff(d2, axis=1)
    return (gap <= rel * np.maximum(d2[:, 1:], 1e-300)).any(axis=1)


def corr_select(warped_v, warped_n, lverts, nn):
    """core/fusion.py:264-274 (identical in core/fusion_dm.py:232-241): best_pt = lverts[nn[0]], best_cost = 1;
    neighbour j replaces it when |dot(n', v' - p_j)| < best_cost.  Returns (best index [m], best_cost [m])."""
    v = np.asarray(warped_v, dtype=np.float64)
    n = np.asarray(warped_n, dtype=np.float64)
    best = np.array(nn[:, 0], dtype=np.int64)
    cost = np.ones(len(v))
    for j in range(nn.shape[1]):
        d = v - lverts[nn[:, j]]                  # float64 - float32 -> float64
        c = np.abs((n[:, 0] * d[:, 0] + n[:, 1] * d[:, 1]) + n[:, 2] * d[:, 2])
        take = c < cost
        cost = np.where(take, c, cost)
        best = np.where(take, nn[:, j], best)
    return best, cost


def unsupported(verts, vert_knn, node_pos, node_w):
    """core/fusion.py:212-215: min_i la.norm(node_i - vert) / dg_w_i >= 1 (float32 norm, float64 quotient)."""
    verts = np.asarray(verts, dtype=np.float32)
    node_pos = np.asarray(node_pos, dtype=np.float32)
    w = np.broadcast_to(np.asarray(node_w, dtype=np.float64), (len(node_pos),))
    mn = np.full(len(verts), np.inf)
    for j in range(vert_knn.shape[1]):
        i = vert_knn[:, j]
        nr = odq.norm3_like_la(node_pos[i] - verts)
        mn = np.minimum(mn, nr.astype(np.float64) / w[i])
    return mn >= 1


def uniform_sample(arr, radius):
    """core/util.py:27-47: take the first remaining candidate, drop every candidate closer than `radius` to it (float64
    norm of float32 data), repeat.  Returns (samples, indices)."""
    c = np.asarray(arr)
    if len(c) == 0:
        return np.zeros((0, 3), c.dtype if c.size else np.float32), np.zeros(0, np.int64)
    c64 = c.astype(np.float64)
    tree = cKDTree(c64)
    alive = np.ones(len(c), bool)
    out = []
    for i in range(len(c)):
        if not alive[i]:
            continue
        out.append(i)
        near = np.array(tree.query_ball_point(c64[i], radius * (1 + 1e-9) + 1e-12), dtype=np.int64)
        d = c64[near] - c64[i]
        nr = np.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2])
        alive[near[nr < radius]] = False
    out = np.array(out, dtype=np.int64)
    return c[out], out
