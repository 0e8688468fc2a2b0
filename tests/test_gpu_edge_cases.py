"""Edge cases of the TSDF path on the GPU (through the C ABI) against the oracle: ragged grids and slabs, every k,
N == k, empty / saturated depth maps, weight saturation at wmax, degenerate blends (zero dual quaternions, weights that
underflow in the reference's exp -> identity fallback, core/fusion.py:544-549), and the class-level error behaviour."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ctx():
    import torch
    from dynamicfusion_body_b200 import engine
    return torch, engine


def _run(sc, res, x0, x1, t0, w0, depths=None, dq=None, node_w=None, wmax=100.0, mode=0, lw="scene"):
    torch, engine = _ctx()
    import scenes
    from oracle import tsdf as ot
    depths = sc.depths if depths is None else depths
    dq = sc.node_dq if dq is None else dq
    node_w = sc.node_w if node_w is None else node_w
    lw = sc.lw if isinstance(lw, str) else lw
    vox, idx, tie = scenes.oracle_knn(res, sc.node_pos, sc.k, x0, x1)
    nw = np.full(sc.n_nodes, np.float32(node_w))
    ov, ow, om, ofr = ot.update_projective(t0.astype(np.float64), w0.astype(np.float64), vox, idx, sc.node_pos, dq, nw, lw, depths,
                                           sc.K, sc.Kinv, sc.tdist, extrinsics=sc.extrinsics, wmax=wmax)
    wf = engine.DeviceWarpField(sc.k)
    wf.set_nodes(sc.node_pos, dq, np.float32(node_w))
    vol = engine.DeviceVolume(res, x0, x1, tsdf=t0, weight=w0)
    m, f = engine.update_projective(vol, wf, lw, torch.from_numpy(np.ascontiguousarray(depths)).cuda(), sc.K, sc.Kinv, sc.extrinsics, sc.tdist,
                                    wmax=wmax, mode=mode, want_masks=True)
    gv, gw = vol.tsdf.cpu().numpy().ravel(), vol.weight.cpu().numpy().ravel()
    ok = ~tie
    gm, gf = m.cpu().numpy(), f.cpu().numpy()
    for v in range(len(depths)):
        assert np.array_equal(((gm >> v) & 1).astype(bool)[ok], om[v][ok])
        assert np.array_equal(((gf >> v) & 1).astype(bool)[ok], ofr[v][ok])
    assert np.abs(gv - ov)[ok].max() <= 1e-5 * sc.tdist
    assert (np.abs(gw - ow) / np.maximum(1, ow))[ok].max() <= 1e-6
    knn = wf.knn_table(res, x0, x1).cpu().numpy().view(np.uint16).astype(np.int64)
    assert np.array_equal(knn[ok], idx[ok])
    return om, vol


@pytest.mark.parametrize("res,slab", [((33, 17, 29), (0, 33)), ((33, 17, 29), (5, 19)), ((9, 70, 131), (2, 9)), ((5, 6, 7), (0, 5))])
def test_ragged_grids_and_slabs(res, slab):
    from dynamicfusion_body_b200 import synth
    import scenes
    sc = synth.make_scene(res=32, k=4, n_nodes=120, seed=3, rows=64, cols=80, background=True)
    n = (slab[1] - slab[0]) * res[1] * res[2]
    t0, w0 = scenes.initial_state(n, tdist=sc.tdist)
    for mode in (0, 1):
        _run(sc, res, slab[0], slab[1], t0, w0, mode=mode)


@pytest.mark.parametrize("k", [1, 2, 3, 5, 6, 7, 8])
def test_every_k(k):
    from dynamicfusion_body_b200 import synth
    import scenes
    sc = synth.make_scene(res=24, k=k, n_nodes=60, seed=k, rows=48, cols=64, background=True)
    t0, w0 = scenes.initial_state(24 ** 3, tdist=sc.tdist)
    _run(sc, (24, 24, 24), 0, 24, t0, w0)


@pytest.mark.parametrize("seed,max_disp,views,k", [(1, 3.0, 1, 4), (2, 10.0, 1, 4), (3, 1.0, 3, 8), (4, 30.0, 1, 4)])
def test_wild_warp_fields_stay_conservative(seed, max_disp, views, k):
    """Large translations and rotating nodes: the warp is far from rigid, the brick and region bounds are loose, the
    fast tier's error margins are at their widest. Results must not change -- only the share of deferred voxels may."""
    from dynamicfusion_body_b200 import synth
    import scenes
    import dataclasses
    R = 64
    sc = synth.make_scene(res=R, k=k, n_nodes=300, seed=seed, rows=120, cols=160, background=True, max_disp=max_disp, n_views=views)
    rng = np.random.default_rng(seed)
    ax = rng.normal(size=(sc.n_nodes, 3))
    ang = np.deg2rad(rng.uniform(0, 3, sc.n_nodes))
    extra = synth.axis_angle_dq(ax, ang, rng.normal(size=(sc.n_nodes, 3)) * max_disp * 0.2).astype(np.float32)
    dq = (sc.node_dq * 0.7 + extra * 0.3).astype(np.float32)
    t0, w0 = scenes.initial_state(R ** 3, tdist=sc.tdist)
    _run(sc, (R, R, R), 0, R, t0, w0, dq=dq)


def test_n_nodes_equals_k():
    """The reference queries k+1 neighbours and drops the last (core/fusion.py:175-176), which needs N >= k+1; with
    N == k every node is a neighbour of every voxel -- the table must still be the distance-sorted permutation."""
    from dynamicfusion_body_b200 import synth
    import dataclasses
    import scenes
    sc = synth.make_scene(res=20, k=4, n_nodes=40, seed=2, rows=48, cols=64, background=True)
    sc = dataclasses.replace(sc, node_pos=sc.node_pos[:4].copy(), node_dq=sc.node_dq[:4].copy(), node_idx=sc.node_idx[:4])
    t0, w0 = scenes.initial_state(20 ** 3, tdist=sc.tdist)
    _run(sc, (20, 20, 20), 0, 20, t0, w0)


def test_empty_and_saturated_depth_and_wmax():
    from dynamicfusion_body_b200 import synth
    import scenes
    sc = synth.make_scene(res=32, k=4, n_nodes=100, seed=5, rows=64, cols=80)
    R = 32
    t0, w0 = scenes.initial_state(R ** 3, tdist=sc.tdist)
    om, vol = _run(sc, (R, R, R), 0, R, t0, w0, depths=np.zeros_like(sc.depths))           # no measurement anywhere
    assert not om.any() and np.array_equal(vol.tsdf.cpu().numpy().ravel(), t0)
    wall = np.full_like(sc.depths, -5.0 * R)                                               # everything is free space
    w_sat = np.full(R ** 3, 99.5, np.float32)
    om, vol = _run(sc, (R, R, R), 0, R, t0, w_sat, depths=wall, wmax=100.0)
    assert om.all() and (vol.weight.cpu().numpy() == 100.0).all()                          # min(w+1, wmax)
    om, vol = _run(sc, (R, R, R), 0, R, t0, np.full(R ** 3, 100.0, np.float32), depths=wall, wmax=100.0)
    assert (vol.weight.cpu().numpy() == 100.0).all()


@pytest.mark.parametrize("mode", [0, 1])
def test_camera_inside_the_volume(mode):
    """The reference never tests the sign of the projective divisor (core/util.py:312-320): a voxel BEHIND the camera is
    mirrored into the image and, with lpos_z < 0, always lands in free space (clamped update); voxel centres exactly on the
    camera plane divide by zero -> `project_to_pixel` returns None -> skipped.  Same here, masks bit for bit."""
    import scenes
    sc = scenes.edge_scene("camera_inside")
    R = 32
    t0, w0 = scenes.initial_state(R ** 3, tdist=sc.tdist)
    with np.errstate(all="ignore"):
        om, vol = _run(sc, (R, R, R), 0, R, t0, w0, mode=mode)
    m = om[0].reshape(R, R, R)
    assert m[:, :, :16].any() and m[:, :, 17:].any() and not m[:, :, 16].any()      # behind: updated; on the camera plane: never


@pytest.mark.parametrize("views", [1, 2])
def test_general_intrinsics(views):
    """K with skew and a third row other than (0,0,1): the brick classifier stands down (it needs a pinhole K), the
    per-voxel tiers follow `project_to_pixel` / K^-1 literally (core/util.py:312-320, core/fusion_dm.py:194-201)."""
    import scenes
    sc = scenes.edge_scene("general_K", views)
    R = 32
    t0, w0 = scenes.initial_state(R ** 3, tdist=sc.tdist)
    om, vol = _run(sc, (R, R, R), 0, R, t0, w0)
    assert om.any() and not om.all()
    assert vol.workspace.stats()["bricks_streamed"] == 0


@pytest.mark.parametrize("mode", [0, 1])
def test_nonfinite_and_positive_depth_pixels(mode):
    """Invalid sensor values follow the reference's comparisons (core/fusion_dm.py:196-203): z = -dm must be > 0, so NaN,
    +inf and positive depth values are skipped; -inf passes that test but K^-1 * (z*u, z*v, z) multiplies it by the zeros
    of a pinhole K^-1, tsdf_l is NaN and `tsdf_l > -tdist` fails: skipped as well."""
    import scenes
    sc = scenes.edge_scene("invalid_depth")
    R = 32
    t0, w0 = scenes.initial_state(R ** 3, tdist=sc.tdist)
    with np.errstate(invalid="ignore"):
        om, vol = _run(sc, (R, R, R), 0, R, t0, w0, mode=mode)
    assert om.any() and not om.all() and np.isfinite(vol.tsdf.cpu().numpy()).all()


def test_degenerate_blends_fall_back_like_the_reference():
    """All-zero node dual quaternions and weights that underflow in float64 both make the blended dq the zero vector;
    the reference then substitutes the identity (core/fusion.py:544-549)."""
    from dynamicfusion_body_b200 import synth
    import scenes
    sc = synth.make_scene(res=32, k=4, n_nodes=100, seed=6, rows=64, cols=80, background=True)
    R = 32
    t0, w0 = scenes.initial_state(R ** 3, tdist=sc.tdist)
    om, _ = _run(sc, (R, R, R), 0, R, t0, w0, dq=np.zeros_like(sc.node_dq))
    assert om.any()
    # dg_w so small that exp(-(d/2w)^2) == 0.0 for most voxels (|arg| > 745), partially underflowing elsewhere
    om, vol = _run(sc, (R, R, R), 0, R, t0, w0, node_w=0.05)
    assert om.any()
    for mode in (0, 1):
        _run(sc, (R, R, R), 0, R, t0, w0, node_w=0.3, mode=mode)
    # non-unit initial transforms of the reference (Q5) with its float32 global dq (Q3 float32 product path)
    dq5 = np.tile(np.array([1, 0, 0, 0, 0, 0.01, 0.01, 0], np.float32), (sc.n_nodes, 1))
    lw32 = (sc.lw + 0).astype(np.float32)
    _run(sc, (R, R, R), 0, R, t0, w0, dq=dq5, lw=lw32)


def test_class_level_errors_and_signatures():
    torch, engine = _ctx()
    from dynamicfusion_body_b200 import Fusion, FusionDM, FusionDM_GPU, synth
    sc = synth.make_scene(res=16, k=4, n_nodes=30, seed=1, rows=32, cols=40)
    fus = Fusion(-sc.tdist, knn=4, use_cnn=False)                       # abs(trunc_distance), core/fusion.py:56
    assert fus._tdist == sc.tdist and fus._lw.dtype == np.float32 and np.allclose(fus._lw, [1, 0, 0, 0, 0, 0.1, 0, 0])
    with pytest.raises(ValueError):
        fus.InitializeCanonicalSpace(tsdf=np.zeros((4, 4)))
    fus.InitializeCanonicalSpace(tsdf=np.full((16, 16, 16), sc.tdist), K=sc.K, vertices=sc.vertices, normals=sc.normals, nodes=sc.nodes_as_reference_tuples())
    with pytest.raises(ValueError, match="tsdf of live frame has not been loaded"):
        fus.updateTSDF()
    with pytest.raises(ValueError):
        fus.updateTSDF(np.zeros((3, 3)))
    with pytest.raises(ValueError):
        fus.fuseFrame(sc.depths, extrinsics=[np.eye(4)[:3]] * 3)
    with pytest.raises(ValueError):
        fus.fuseFrame(np.zeros((9, 8, 8), np.float32))                  # more views than one pass takes
    nodes = fus._nodes
    assert len(nodes) == sc.n_nodes and nodes[0][2].shape == (8,) and isinstance(nodes[0][3], float)
    assert fus.knn_indices().shape == (16 ** 3, 4)
    p = fus.warp(sc.vertices[0], m_lw=fus._lw)
    assert p.shape == (3,)
    assert fus.dq_blend(sc.vertices[0]).shape == (8,)
    fus.surface_extractor = None                                        # extraction switched off: the callers say so
    for name in ("update_graph", "marching_cubes"):
        with pytest.raises(NotImplementedError):
            getattr(fus, name)()
    fdm = FusionDM_GPU(0.2, sc.K, tsdf_res=16)
    assert isinstance(fdm, FusionDM) and fdm._tsdf.shape == (16, 16, 16) and (fdm._tsdf == np.float32(0.2)).all() and (fdm._tsdfw == 0).all()
    fdm.verbose_gpu()                                                   # core/fusion_dm.py:576 (device listing)
    assert fdm.write_live_frame_mesh("", "", "") is None                # core/fusion_dm.py:357-358 (empty in the reference)
    with pytest.raises(ValueError):
        fdm.compute_live_tsdf([sc.depths[0]], [])
    t, w = fdm.fuseDepths(sc.depths[0], np.eye(4)[:3], np.full((16, 16, 16), 0.2), np.zeros((16, 16, 16)))
    assert t.shape == (16, 16, 16) and t.dtype == np.float64


@pytest.mark.parametrize("views,k", [(1, 4), (3, 8)])
def test_deferred_list_overflow_is_still_exact(views, k):
    """More voxels deferred than the work list holds: the rest is marked in the overflow bitmap and swept by the exact
    pass (brick-level decisions included) -- same result as with a list that fits."""
    torch, engine = _ctx()
    from dynamicfusion_body_b200 import synth
    import scenes
    from oracle import tsdf as ot
    sc = synth.make_scene(res=40, k=k, n_nodes=150, seed=8, rows=96, cols=128, n_views=views)
    R = 40
    res = (R, R, R)
    vox, idx, tie = scenes.oracle_knn(res, sc.node_pos, sc.k, 0, R)
    t0, w0 = scenes.initial_state(R ** 3, tdist=sc.tdist)
    nw = np.full(sc.n_nodes, np.float32(sc.node_w))
    ov, ow, om, ofr = ot.update_projective(t0.astype(np.float64), w0.astype(np.float64), vox, idx, sc.node_pos, sc.node_dq, nw, sc.lw,
                                           sc.depths, sc.K, sc.Kinv, sc.tdist, extrinsics=sc.extrinsics)
    wf = engine.DeviceWarpField(sc.k)
    wf.set_nodes(sc.node_pos, sc.node_dq, np.float32(sc.node_w))
    depths = torch.from_numpy(sc.depths).cuda()
    outs = []
    for capacity in (None, 1024):
        vol = engine.DeviceVolume(res, tsdf=t0, weight=w0)
        if capacity:
            vol.workspace.capacity = capacity
        for frame in range(2):                           # twice: the bitmap must come back clean
            m, f = engine.update_projective(vol, wf, sc.lw, depths, sc.K, sc.Kinv, sc.extrinsics, sc.tdist, want_masks=True)
        st = vol.workspace.stats()
        assert st["exact_processed"] == st["deferred"] and (capacity is None or st["deferred"] > capacity)
        assert not vol.workspace.overflow_bits.any()
        outs.append((vol.tsdf.cpu().numpy().ravel(), vol.weight.cpu().numpy().ravel(), m.cpu().numpy()))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1]) and np.array_equal(outs[0][2], outs[1][2])
    ok = ~tie
    for v in range(views):
        assert np.array_equal(((outs[1][2] >> v) & 1).astype(bool)[ok], om[v][ok])
    # a1 through the same mechanism (live TSDF of the node-warped mesh, no global lw, a thick band)
    wv = synth.blend_warp(sc.vertices, sc.node_pos, sc.node_dq, nw, sc.vert_knn, lw=None)
    td = 3.0
    live = torch.from_numpy(np.clip(synth.mesh_sdf_volume(res, wv, sc.normals), -1.5 * td, 1.5 * td)).cuda()
    a1 = []
    for capacity in (None, 128):
        vol = engine.DeviceVolume(res, tsdf=t0, weight=w0)
        if capacity:
            vol.workspace.capacity = capacity
        engine.update_volume(vol, wf, None, live, td)
        st = vol.workspace.stats()
        assert st["exact_processed"] == st["deferred"] > 128
        assert not vol.workspace.overflow_bits.any()
        a1.append((vol.tsdf.cpu().numpy(), vol.weight.cpu().numpy()))
    assert np.array_equal(a1[0][0], a1[1][0]) and np.array_equal(a1[0][1], a1[1][1])
