"""CPU oracle drivers used by bench.py's cpu_baseline / `--impl reference` legs and by smoke() --
TEST INFRASTRUCTURE ONLY (see oracle/dq.py header).  They time / evaluate the numpy restatement of the
reference path on a bounded sample of voxels of the benchmark workload.

The kNN here uses scipy's KD-tree exactly like the reference (core/fusion.py:119,175)."""
import time

import numpy as np
from scipy.spatial import cKDTree

from . import tsdf as ot


def sample_voxels(res, n, seed=0, x0=0, x1=None):
    """n distinct random voxel multi-indices (float32, like np.array(it.multi_index, dtype=np.float32))."""
    x1 = res[0] if x1 is None else x1
    rng = np.random.default_rng(seed)
    total = (x1 - x0) * res[1] * res[2]
    n = min(n, total)
    lin = rng.choice(total, size=n, replace=False) if total < (1 << 26) else np.unique(rng.integers(0, total, size=n))
    x, rem = np.divmod(lin, res[1] * res[2])
    y, z = np.divmod(rem, res[2])
    return lin, np.stack([x + x0, y, z], 1).astype(np.float32)


def projective_on_sample(scene, vox, tsdf, tsdfw, tree=None):
    """Reference-path result (a3) for voxels `vox`: KD-tree kNN + warp + project + fuse."""
    tree = tree if tree is not None else cKDTree(scene["node_pos"].astype(np.float64))
    _, idx = tree.query(vox.astype(np.float64), k=scene["k"] + 1)
    idx = idx[:, :-1]
    return ot.update_projective(tsdf, tsdfw, vox, idx, scene["node_pos"], scene["node_dq"], scene["node_w"], scene["lw"],
                                scene["depths"], scene["K"], scene["Kinv"], scene["tdist"],
                                extrinsics=scene.get("extrinsics"))


def time_projective(scene, res, n_sample, seed=0, chunk=200_000):
    """voxels/s of the oracle on one core over a sample of `n_sample` voxels of the (res) grid."""
    lin, vox = sample_voxels(res, n_sample, seed)
    tree = cKDTree(scene["node_pos"].astype(np.float64))
    t0 = time.perf_counter()
    for s in range(0, len(vox), chunk):
        v = vox[s:s + chunk]
        projective_on_sample(scene, v, np.full(len(v), scene["tdist"]), np.zeros(len(v)), tree)
    dt = time.perf_counter() - t0
    return len(vox) / dt, len(vox), dt


def _worker(args):
    scene, res, n, seed = args
    return time_projective(scene, res, n, seed)


def time_projective_parallel(scene, res, n_sample, procs):
    """Same sample split over `procs` worker processes (x-slab style embarrassingly parallel split);
    returns aggregate voxels/s measured as total voxels / wall time of the slowest worker."""
    import multiprocessing as mp
    per = max(1, n_sample // procs)
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(procs) as pool:
        out = pool.map(_worker, [(scene, res, per, 100 + i) for i in range(procs)])
    wall = time.perf_counter() - t0
    nvox = sum(o[1] for o in out)
    return nvox / max(o[2] for o in out), nvox, wall


def scene_dict(sc):
    """Plain-dict view of a dynamicfusion_body_b200.synth.Scene for the functions above."""
    return {"node_pos": sc.node_pos, "node_dq": sc.node_dq, "node_w": np.full(sc.n_nodes, sc.node_w), "lw": sc.lw,
            "depths": sc.depths, "K": sc.K, "Kinv": sc.Kinv, "tdist": sc.tdist, "extrinsics": sc.extrinsics, "k": sc.k}
