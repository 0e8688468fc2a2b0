"""The kernels' per-voxel / per-residual logic (csrc/dfb_math.h, dfb_voxel.h, dfb_gn.h), compiled for the host by
tests/hostshim, against the oracle -- so the two arithmetic tiers are checked on the CPU box as well.  The same
comparisons run against the real CUDA kernels in tests/test_gpu_*.py."""
import numpy as np
import pytest

import hostshim_api as hs
import scenes
from dynamicfusion_body_b200 import synth
from oracle import dq as odq
from oracle import gn as ogn
from oracle import tsdf as ot


@pytest.fixture(scope="module")
def sc():
    return synth.make_scene(res=40, k=4, n_nodes=200, seed=1, rows=96, cols=128)


@pytest.mark.parametrize("mode", [0, 1])
def test_projective_tiers(sc, mode):
    R = sc.res
    vox, idx, tie = scenes.oracle_knn((R, R, R), sc.node_pos, sc.k)
    t0, w0 = scenes.initial_state(R ** 3, tdist=sc.tdist)
    nw = np.full(sc.n_nodes, sc.node_w)
    ov, ow, om, ofr = ot.update_projective(t0.astype(np.float64), w0.astype(np.float64), vox, idx, sc.node_pos, sc.node_dq, nw, sc.lw,
                                           sc.depths, sc.K, sc.Kinv, sc.tdist)
    wf = hs.HostWarpField(sc.node_pos, sc.node_dq, np.float32(sc.node_w), sc.k, knn=idx, lw=sc.lw)
    tv, tw = t0.copy(), w0.copy()
    mask, frus, cls, nunc = hs.update_projective(tv, tw, (R, R, R), wf, sc.depths, sc.K, sc.Kinv, sc.tdist, mode=mode)
    assert np.array_equal(scenes.bits(mask, 0), om[0]) and np.array_equal(scenes.bits(frus, 0), ofr[0])
    assert np.abs(tv - ov).max() <= 1e-5 * sc.tdist
    assert (np.abs(tw - ow) / np.maximum(1, ow)).max() <= 1e-6
    if mode == 0:
        assert nunc < 0.3 * R ** 3          # the fast tier must decide the bulk of the volume
        # every voxel the fast tier decided must agree with the oracle's classification
        assert not (om[0] & (cls == 0)).any()


def test_projective_multiview_extrinsics():
    s4 = synth.make_scene(res=32, k=8, n_nodes=120, seed=4, n_views=3, rows=64, cols=80)
    R = s4.res
    vox, idx, tie = scenes.oracle_knn((R, R, R), s4.node_pos, s4.k)
    t0, w0 = scenes.initial_state(R ** 3, tdist=s4.tdist)
    nw = np.full(s4.n_nodes, s4.node_w)
    ov, ow, om, ofr = ot.update_projective(t0.astype(np.float64), w0.astype(np.float64), vox, idx, s4.node_pos, s4.node_dq, nw, s4.lw,
                                           s4.depths, s4.K, s4.Kinv, s4.tdist, extrinsics=s4.extrinsics)
    wf = hs.HostWarpField(s4.node_pos, s4.node_dq, np.float32(s4.node_w), s4.k, knn=idx, lw=s4.lw)
    tv, tw = t0.copy(), w0.copy()
    mask, frus, cls, nunc = hs.update_projective(tv, tw, (R, R, R), wf, s4.depths, s4.K, s4.Kinv, s4.tdist, extrinsics=s4.extrinsics)
    for v in range(3):
        assert np.array_equal(scenes.bits(mask, v), om[v]) and np.array_equal(scenes.bits(frus, v), ofr[v])
    assert np.abs(tv - ov).max() <= 1e-5 * s4.tdist


@pytest.mark.parametrize("lw", [np.array([1, 0, 0, 0, 0, 0.1, 0, 0], np.float32), np.array([1, 0, 0, 0, 0, 0.1, 0, 0.05]), None])
def test_volume_update_tiers(sc, lw):
    R = sc.res
    vox, idx, tie = scenes.oracle_knn((R, R, R), sc.node_pos, sc.k)
    nw = np.full(sc.n_nodes, sc.node_w)
    wv = synth.blend_warp(sc.vertices, sc.node_pos, sc.node_dq, nw, sc.vert_knn, lw=None if lw is None else lw.astype(np.float64))
    live = synth.mesh_sdf_volume((R + 2, R, R + 1), wv, sc.warped_normals)
    for tdist, lv in ((float(live.max()), live), (2.0, np.clip(live, -3.0, 3.0))):
        t0, w0 = scenes.initial_state(R ** 3, tdist=tdist)
        ov, ow, om = ot.update_volume(t0.astype(np.float64), w0.astype(np.float64), lv, vox, idx, sc.node_pos, sc.node_dq, nw, lw, tdist)
        wf = hs.HostWarpField(sc.node_pos, sc.node_dq, np.float32(sc.node_w), sc.k, knn=idx, lw=lw)
        for mode in (0, 1):
            tv, tw = t0.copy(), w0.copy()
            mask, cls, nunc = hs.update_volume(tv, tw, (R, R, R), wf, lv, tdist, mode=mode)
            assert np.array_equal(mask.astype(bool), om)
            assert np.abs(tv - ov).max() <= 1e-5 * tdist
            assert (np.abs(tw - ow) / np.maximum(1, ow)).max() <= 1e-6


def test_rigid_paths(sc):
    R = sc.res
    vox = ot.voxel_grid((R, R, R))
    K = np.array([[200., 0, 80], [0, 200, 60], [0, 0, 1]]); Kinv = np.linalg.inv(K)
    ang = 0.1
    Rm = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
    scale = 12 * 1.3 / R; center = np.array([-0.03, -0.43, -5.6])
    lw34 = np.concatenate([Rm, -Rm @ center[:, None] + np.array([[0.1], [0.05], [22.0]])], 1)
    vw = scale * (sc.vertices.astype(np.float64) - R / 2) + center
    dm = synth.render_depth(vw @ lw34[:, :3].T + lw34[:, 3], sc.faces, K, 120, 160)
    t0, w0 = scenes.initial_state(R ** 3, tdist=0.2)
    ov, ow, om, ofr = ot.fuse_depth_rigid(t0.astype(np.float64), w0.astype(np.float64), vox, dm, lw34, K, Kinv, 0.2, R, scale=scale, center=center)
    assert om.mean() > 0.02
    for mode in (0, 1):
        tv, tw = t0.copy(), w0.copy()
        mask, frus, cls, nunc = hs.fuse_depth_rigid(tv, tw, (R, R, R), R, dm, lw34, K, Kinv, scale, center, 0.2, mode=mode)
        assert np.array_equal(mask.astype(bool), om) and np.array_equal(frus.astype(bool), ofr)
        assert np.abs(tv - ov).max() <= 1e-5 * 0.2 and np.array_equal(tw, ow.astype(np.float32))
    # FusionDM.updateTSDF
    lw = np.array([1, 0, 0, 0, 0, 0.1, 0, 0], np.float32)
    live = synth.mesh_sdf_volume((R, R, R), sc.vertices, sc.normals)
    ov, ow, om = ot.update_rigid_volume(t0.astype(np.float64), w0.astype(np.float64), live, vox, lw, 4.0)
    wf0 = hs.HostWarpField(np.zeros((0, 3)), np.zeros((0, 8)), np.zeros(0), 0, lw=lw)
    tv, tw = t0.copy(), w0.copy()
    mask, cls, nunc = hs.update_volume(tv, tw, (R, R, R), wf0, live, 4.0)
    assert np.array_equal(mask.astype(bool), om) and np.abs(tv - ov).max() <= 1e-5 * 4.0


def test_warp_points(sc):
    nwk = np.full((len(sc.vertices), sc.k), sc.node_w)
    for lw in (np.array([1, 0, 0, 0, 0, 0.1, 0, 0], np.float32), sc.lw, None):
        wf = hs.HostWarpField(sc.node_pos, sc.node_dq, np.float32(sc.node_w), sc.k, lw=lw)
        p, n = hs.warp_points(sc.vertices, sc.normals, sc.vert_knn, wf)
        op, on = odq.warp(sc.vertices, sc.node_pos[sc.vert_knn], sc.node_dq[sc.vert_knn], nwk, lw=lw, normal=sc.normals)
        assert np.abs(p - op).max() <= 1e-7 * max(1.0, np.abs(op).max()) and np.abs(n - on).max() <= 1e-7


@pytest.mark.parametrize("huber", [False, True])
def test_gn_residuals_and_normal_equations(sc, huber):
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(0)
    sel = rng.choice(len(sc.vertices), 300, replace=False)
    verts, norms, knn = sc.vertices[sel], sc.normals[sel], sc.vert_knn[sel]
    corr = sc.warped_vertices[sel] + rng.normal(size=(300, 3)) * 0.3
    _, nvi = cKDTree(verts.astype(np.float64)).query(sc.node_pos.astype(np.float64))
    nodes = slice(0, sc.n_nodes)
    nw32 = np.float32(sc.node_w)
    N = sc.n_nodes
    for lw in (np.array([1, 0, 0, 0, 0, 0.1, 0, 0], np.float32), sc.lw):
        Gp = hs.HostGN(verts, norms, corr, knn, sc.node_pos, nw32, nvi, lw, 0.5, huber=huber, f_scale=0.2)
        for xdt in (np.float32, np.float64):
            x = sc.node_dq.reshape(-1).astype(xdt)
            fo = ogn.computef(x, verts, norms, corr, knn, sc.node_pos, float(nw32), nvi, lw, 0.5)
            assert np.abs(fo - Gp.residuals(x)).max() <= 1e-12 * max(1, np.abs(fo).max())
        x = sc.node_dq.reshape(-1).astype(np.float64) + rng.normal(size=8 * N) * 1e-3
        J, r = ogn.jacobian(x, verts, norms, corr, knn, sc.node_pos, float(nw32), nvi, lw, 0.5)
        Ho, go = ogn.normal_equations(J, r, f_scale=0.2, huber=huber)
        H, g, cost = Gp.normal_eq_dense(x)
        assert np.abs(H - Ho).max() <= 1e-10 * np.abs(Ho).max() and np.abs(g - go).max() <= 1e-10 * np.abs(go).max()
        assert abs(cost[0] - ogn.robust_cost(r, 0.2, huber)) <= 1e-10 * cost[0]
        Jl, rl = ogn.lw_jacobian(lw, verts, norms, corr, knn, sc.node_pos, x.reshape(-1, 8), float(nw32))
        wl = ogn.huber_weights(rl, 0.2, huber)
        Hl, gl, cl = Gp.lw_normal_eq(x.reshape(-1, 8), lw)
        assert np.abs(Jl.T @ (wl[:, None] * Jl) - Hl).max() <= 1e-10 * np.abs(Hl).max()


def test_brick_culling_is_conservative():
    """dfb_brick.h: a brick classified CLAMP/SKIP must agree with the oracle for every one of its voxels, and the final
    volume must be identical to the un-culled result."""
    s = synth.make_scene(res=128, k=4, n_nodes=600, seed=3, background=True, rows=240, cols=320)
    R = s.res
    from scipy.spatial import cKDTree
    vox = ot.voxel_grid((R, R, R))
    _, idx = cKDTree(s.node_pos.astype(np.float64)).query(vox.astype(np.float64), k=4)
    t0, w0 = scenes.initial_state(R ** 3, tdist=s.tdist)
    nw = np.full(s.n_nodes, s.node_w)
    ov, ow, om, ofr = ot.update_projective(t0.astype(np.float64), w0.astype(np.float64), vox, idx, s.node_pos, s.node_dq, nw, s.lw,
                                           s.depths, s.K, s.Kinv, s.tdist)
    wf = hs.HostWarpField(s.node_pos, s.node_dq, np.float32(s.node_w), 4, knn=idx, lw=s.lw)
    res = []
    for bricks in (False, True):
        cv = hs.set_bricks(idx, 4, (R, R, R), enable=bricks, regions=True)
        tv, tw = t0.copy(), w0.copy()
        mask, frus, cls, nunc = hs.update_projective(tv, tw, (R, R, R), wf, s.depths, s.K, s.Kinv, s.tdist)
        res.append((tv, tw, mask, frus))
        assert np.array_equal(scenes.bits(mask, 0), om[0]) and np.array_equal(scenes.bits(frus, 0), ofr[0])
    nreg = (R // 16) * (R // 16) * (R // 32)
    assert hs.regions_resolved() > 0.2 * nreg            # whole 16x16x32 regions are settled by one box test
    hs.set_bricks(enable=False)
    # masks and weights are identical; values may differ by an fp32 ulp where one run resolved a voxel in the fp32
    # tier and the other in the float64 tier
    for a, b in zip(res[0][1:], res[1][1:]):
        assert np.array_equal(a, b)
    assert np.array_equal(res[0][0], res[1][0])
    assert (cv != 255).mean() > 0.3                      # a sizeable part of the volume never needs the per-voxel tier
    assert not (om[0] & (cv == 0)).any()                 # SKIP bricks contain no updated voxel
    assert om[0][(cv != 0) & (cv != 255)].all()          # CLAMP bricks contain only updated voxels


@pytest.mark.parametrize("axis", ["x", "y"])
def test_brick_culling_camera_across_the_bricks(axis):
    """The same conservativeness with the camera along x / y (a 4x4x32 brick then spans 32 voxels across the image): masks and
    frustum bits equal the oracle's with bricks, regions and the quad pre-test on; SKIP / CLAMP bricks hold what they claim."""
    s = synth.make_scene(res=64, k=4, n_nodes=300, seed=2, background=True, rows=96, cols=128, view_axis=axis)
    R = s.res
    from scipy.spatial import cKDTree
    vox = ot.voxel_grid((R, R, R))
    _, idx = cKDTree(s.node_pos.astype(np.float64)).query(vox.astype(np.float64), k=4)
    t0, w0 = scenes.initial_state(R ** 3, tdist=s.tdist)
    nw = np.full(s.n_nodes, s.node_w)
    ov, ow, om, ofr = ot.update_projective(t0.astype(np.float64), w0.astype(np.float64), vox, idx, s.node_pos, s.node_dq, nw, s.lw,
                                           s.depths, s.K, s.Kinv, s.tdist)
    wf = hs.HostWarpField(s.node_pos, s.node_dq, np.float32(s.node_w), 4, knn=idx, lw=s.lw)
    cv = hs.set_bricks(idx, 4, (R, R, R), enable=True, regions=True)
    tv, tw = t0.copy(), w0.copy()
    mask, frus, cls, nunc = hs.update_projective(tv, tw, (R, R, R), wf, s.depths, s.K, s.Kinv, s.tdist)
    hs.set_bricks(enable=False)
    assert np.array_equal(scenes.bits(mask, 0), om[0]) and np.array_equal(scenes.bits(frus, 0), ofr[0])
    assert np.abs(tv - ov).max() <= 1e-5 * s.tdist
    assert (cv != 255).any()                              # (at 64^3 a 32-voxel brick is half the grid: most bricks are MIXED on any axis)
    assert not (om[0] & (cv == 0)).any() and om[0][(cv != 0) & (cv != 255)].all()


@pytest.mark.parametrize("case", ["zero_dq", "w_tiny", "w_small", "q5_lw32"])
def test_degenerate_blends(case):
    """Blend weights that vanish in the reference: `w * dg_dq` is a float32 product (numpy weak scalar), so weights below the
    float32 range drop out and an all-zero blend becomes the identity (core/fusion.py:544-549).  The fp32 tier must hand
    exactly those voxels to the exact tier."""
    s = synth.make_scene(res=32, k=4, n_nodes=100, seed=6, rows=64, cols=80, background=True)
    R = 32
    t0, w0 = scenes.initial_state(R ** 3, tdist=s.tdist)
    vox, idx, tie = scenes.oracle_knn((R, R, R), s.node_pos, 4)
    dq, node_w, lw = s.node_dq, s.node_w, s.lw
    if case == "zero_dq":
        dq = np.zeros_like(s.node_dq)
    elif case == "w_tiny":
        node_w = 0.05
    elif case == "w_small":
        node_w = 0.3
    else:
        dq = np.tile(np.array([1, 0, 0, 0, 0, 0.01, 0.01, 0], np.float32), (s.n_nodes, 1))
        lw = s.lw.astype(np.float32)
    nw = np.full(s.n_nodes, np.float32(node_w))
    ov, ow, om, ofr = ot.update_projective(t0.astype(np.float64), w0.astype(np.float64), vox, idx, s.node_pos, dq, nw, lw, s.depths, s.K,
                                           s.Kinv, s.tdist)
    wf = hs.HostWarpField(s.node_pos, dq, np.float32(node_w), 4, knn=idx, lw=lw)
    for mode, bricks in ((0, False), (0, True), (1, False)):
        hs.set_bricks(idx, 4, (R, R, R), enable=bricks)
        tv, tw = t0.copy(), w0.copy()
        mask, frus, cls, nunc = hs.update_projective(tv, tw, (R, R, R), wf, s.depths, s.K, s.Kinv, s.tdist, mode=mode)
        assert np.array_equal(scenes.bits(mask, 0), om[0]) and np.array_equal(scenes.bits(frus, 0), ofr[0])
        assert np.abs(tv - ov).max() <= 1e-5 * s.tdist
    hs.set_bricks(enable=False)


@pytest.mark.parametrize("kind", ["invalid_depth", "general_K", "camera_inside"])
@pytest.mark.parametrize("bricks", [False, True])
def test_projective_edge_scenes(kind, bricks):
    """NaN / +-inf / positive depth pixels, a non-pinhole K, voxels behind and on the camera plane: the host build of the
    two tiers (with and without the brick / region classifier in front) against the oracle, masks bit for bit."""
    s = scenes.edge_scene(kind)
    R = s.res
    vox, idx, tie = scenes.oracle_knn((R, R, R), s.node_pos, s.k)
    t0, w0 = scenes.initial_state(R ** 3, tdist=s.tdist)
    nw = np.full(s.n_nodes, s.node_w)
    with np.errstate(all="ignore"):
        ov, ow, om, ofr = ot.update_projective(t0.astype(np.float64), w0.astype(np.float64), vox, idx, s.node_pos, s.node_dq, nw, s.lw,
                                               s.depths, s.K, s.Kinv, s.tdist, extrinsics=s.extrinsics)
    wf = hs.HostWarpField(s.node_pos, s.node_dq, np.float32(s.node_w), s.k, knn=idx, lw=s.lw)
    cls_b = hs.set_bricks(idx, s.k, (R, R, R), enable=True) if bricks else hs.set_bricks(enable=False)
    try:
        tv, tw = t0.copy(), w0.copy()
        mask, frus, cls, nunc = hs.update_projective(tv, tw, (R, R, R), wf, s.depths, s.K, s.Kinv, s.tdist, extrinsics=s.extrinsics)
    finally:
        hs.set_bricks(enable=False)
    ok = ~tie
    assert np.array_equal(scenes.bits(mask, 0)[ok], om[0][ok]) and np.array_equal(scenes.bits(frus, 0)[ok], ofr[0][ok])
    assert np.abs(tv - ov)[ok].max() <= 1e-5 * s.tdist
    assert (np.abs(tw - ow) / np.maximum(1, ow))[ok].max() <= 1e-6
    assert om[0].any() and not om[0].all()
