"""Scratch fuzzer (CPU only): random a3 scenes through the HOST build of the kernel logic (tests/hostshim: both arithmetic
tiers + the brick / region classifier) against the oracle.  Masks and frustum bits must agree bit for bit, values to
1e-5 tdist.  usage: python scripts/fuzz_hostshim.py [n_cases] [first_seed]"""
import sys, os, copy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import hostshim_api as hs
import scenes
from dynamicfusion_body_b200 import synth
from oracle import tsdf as ot

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 20
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 100
bad = 0
for seed in range(seed0, seed0 + n_cases):
    rng = np.random.default_rng(seed)
    R = int(rng.choice([20, 24, 32, 40]))
    k = int(rng.choice([1, 2, 3, 4, 8]))
    views = int(rng.choice([1, 1, 2, 3]))
    max_disp = float(rng.choice([0.1, 0.5, 2.0, 8.0]))
    sc = copy.copy(synth.make_scene(res=R, k=k, n_nodes=int(rng.integers(max(k + 1, 20), 160)), seed=seed, rows=int(rng.choice([48, 96])),
                                    cols=int(rng.choice([64, 128])), n_views=views, max_disp=max_disp, background=bool(rng.integers(2)),
                                    cam_dist=float(rng.choice([0.9, 1.7, 3.0])), unit_init=bool(rng.random() < 0.15),
                                    lw_dtype=np.float32 if rng.random() < 0.3 else np.float64, view_axis=str(rng.choice(["z", "z", "x", "y"]))))
    if rng.random() < 0.4:                                            # rotating nodes on top of the smooth field
        ax = rng.normal(size=(sc.n_nodes, 3)); ang = np.deg2rad(rng.uniform(0, 6, sc.n_nodes))
        extra = synth.axis_angle_dq(ax, ang, rng.normal(size=(sc.n_nodes, 3)) * max_disp * 0.3).astype(np.float32)
        sc.node_dq = (sc.node_dq * 0.6 + extra * 0.4).astype(np.float32)
    if rng.random() < 0.3:                                            # non-unit transforms (Q5-like)
        sc.node_dq = (sc.node_dq * rng.uniform(0.5, 1.5, (sc.n_nodes, 1))).astype(np.float32)
    d = sc.depths.copy()
    if rng.random() < 0.3:
        r = rng.random(d.shape); d[r < 0.05] = 0; d[(r > 0.05) & (r < 0.08)] = np.nan; d[(r > 0.08) & (r < 0.1)] = -np.inf
    sc.depths = d
    tdist = float(sc.tdist * rng.choice([0.3, 1.0, 3.0]))
    res = (R, int(R + rng.integers(-3, 4)), int(R + rng.integers(-5, 6)))
    vox, idx, tie = scenes.oracle_knn(res, sc.node_pos, sc.k)
    n = vox.shape[0]
    t0, w0 = scenes.initial_state(n, seed=seed, fresh=bool(rng.random() < 0.3), tdist=tdist)
    nw = np.full(sc.n_nodes, sc.node_w)
    wmax = float(rng.choice([100.0, 5.0]))
    with np.errstate(all="ignore"):
        ov, ow, om, ofr = ot.update_projective(t0.astype(np.float64), w0.astype(np.float64), vox, idx, sc.node_pos, sc.node_dq, nw, sc.lw,
                                               sc.depths, sc.K, sc.Kinv, tdist, extrinsics=sc.extrinsics, wmax=wmax)
    wf = hs.HostWarpField(sc.node_pos, sc.node_dq, np.float32(sc.node_w), sc.k, knn=idx, lw=sc.lw)
    for bricks in (False, True):
        if bricks:
            hs.set_bricks(idx, sc.k, res, enable=True, regions=bool(rng.integers(2)))
        else:
            hs.set_bricks(enable=False)
        tv, tw = t0.copy(), w0.copy()
        mask, frus, cls, nunc = hs.update_projective(tv, tw, res, wf, sc.depths, sc.K, sc.Kinv, tdist, extrinsics=sc.extrinsics, wmax=wmax)
        hs.set_bricks(enable=False)
        ok = ~tie
        errs = []
        for v in range(views):
            if not np.array_equal(scenes.bits(mask, v)[ok], om[v][ok]): errs.append("mask v%d: %d" % (v, (scenes.bits(mask, v)[ok] != om[v][ok]).sum()))
            if not np.array_equal(scenes.bits(frus, v)[ok], ofr[v][ok]): errs.append("frustum v%d: %d" % (v, (scenes.bits(frus, v)[ok] != ofr[v][ok]).sum()))
        dv = np.abs(tv - ov)[ok].max() / tdist
        dw = (np.abs(tw - ow) / np.maximum(1, ow))[ok].max()
        if dv > 1e-5: errs.append("dTSDF %.2e tdist" % dv)
        if dw > 1e-6: errs.append("dW %.2e" % dw)
        print("seed %d R=%s k=%d views=%d disp=%.1f bricks=%d: updated %.3f deferred %.3f  %s" % (seed, res, k, views, max_disp, bricks, om.any(0).mean(), nunc / n, "OK" if not errs else "MISMATCH " + "; ".join(errs)), flush=True)
        bad += bool(errs)
print("cases with a mismatch:", bad)
