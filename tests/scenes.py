"""Shared helpers for the parity tests: seeded volume state and oracle drivers."""
import numpy as np

from oracle import dq as odq
from oracle import tsdf as ot


def initial_state(n, seed=3, fresh=False, tdist=1.0):
    """float32 (tsdf, weight): either the reference's fresh state (tsdf=+tdist, w=0;
    core/fusion_dm.py:61-62) or a random mid-sequence state with ~half zero weights."""
    rng = np.random.default_rng(seed)
    if fresh:
        return np.full(n, tdist, np.float32), np.zeros(n, np.float32)
    t = (rng.normal(size=n) * 0.4 * tdist).clip(-tdist, tdist).astype(np.float32)
    w = np.where(rng.random(n) < 0.5, 0, rng.integers(1, 120, size=n)).astype(np.float32)
    return t, w


def oracle_knn(res, node_pos, k, x0=0, x1=None):
    vox = ot.voxel_grid(res, x0, x1)
    idx, d2 = odq.knn_bruteforce(vox, node_pos, k)
    return vox, idx, odq.knn_has_tie(d2)


def bits(mask_bits, view):
    return ((np.asarray(mask_bits) >> view) & 1).astype(bool)


def edge_scene(kind, views=1):
    """Small a3 scenes for the edge cases that both the GPU tests and the host build of the kernel logic run:
    'invalid_depth' (NaN, +-inf, positive pixels), 'general_K' (skewed K, third row != (0,0,1)),
    'camera_inside' (voxels behind the camera and exactly on the camera plane)."""
    import copy
    from dynamicfusion_body_b200 import synth
    if kind == "invalid_depth":
        sc = copy.copy(synth.make_scene(res=32, k=4, n_nodes=100, seed=6, rows=64, cols=80))
        rng = np.random.default_rng(0)
        d = sc.depths.copy()
        r = rng.random(d.shape)
        d[r < 0.10] = np.nan
        d[(r >= 0.10) & (r < 0.15)] = np.inf
        d[(r >= 0.15) & (r < 0.20)] = -np.inf
        d[(r >= 0.20) & (r < 0.25)] = 7.5
        d[:, 10:30, 20:50] = np.nan                                                        # a whole block without data
        sc.depths = d
    elif kind == "general_K":
        sc = copy.copy(synth.make_scene(res=32, k=4, n_nodes=100, seed=7, rows=64, cols=80, n_views=views))
        K = sc.K.copy()
        K[0, 1] = 1.5; K[1, 0] = -0.7; K[2] = [2e-4, -1e-4, 1.05]
        sc.K, sc.Kinv = K, np.linalg.inv(K)
    elif kind == "camera_inside":
        sc = copy.copy(synth.make_scene(res=32, k=4, n_nodes=100, seed=9, rows=64, cols=80))
        E = np.concatenate([np.eye(3), -np.array([[15.5], [16.25], [16.0]])], 1)           # camera at the volume centre, looking along +z
        sc.extrinsics = E[None]
        sc.lw = np.array([1.0, 0, 0, 0, 0, 0, 0, 0])
        sc.node_dq = np.tile(np.array([1, 0, 0, 0, 0, 0, 0, 0], np.float32), (sc.n_nodes, 1))  # identity warp: voxel z = 16 is the camera plane
        sc.K = np.array([[30.0, 0, 40.0], [0, 30.0, 32.0], [0, 0, 1.0]])
        sc.Kinv = np.linalg.inv(sc.K)
        rng = np.random.default_rng(1)
        sc.depths = -(4.0 + 6.0 * rng.random((1, 64, 80))).astype(np.float32)
    else:
        raise ValueError(kind)
    return sc
