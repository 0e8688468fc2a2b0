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
