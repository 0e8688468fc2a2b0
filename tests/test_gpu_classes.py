"""Class-level drop-in parity (GPU): the reference-named methods of Fusion / FusionDM / FusionDM_GPU against the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_fusiondm_compute_live_tsdf_matches_reference_loop():
    """FusionDM_GPU.compute_live_tsdf (core/fusion_dm.py:95-178, plain branch): fresh volume, every depth map fused with
    scale = 12*std/R, center = avg, hard-coded avg/std (:106-107); numpy in / numpy out."""
    import torch
    from dynamicfusion_body_b200 import FusionDM_GPU, synth
    from oracle import tsdf as ot
    R = 48
    K = np.array([[400., 0, 79.5], [0, 400., 59.5], [0, 0, 1]])
    v64, nrm, faces = synth.load_body_mesh()
    std, avg = 1.3, np.array([-0.03, -0.43, -5.6], dtype='float32')
    scale = 12 * std / R
    verts_w = scale * (v64.astype(np.float64) * (R - 1) / 64.0 - R / 2) + avg.astype(np.float64)
    depths, lws = [], []
    for ang in (0.0, 0.5, -0.4):
        Rm = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
        t = -Rm @ avg.astype(np.float64) + np.array([0.0, 0.0, 30.0])
        lw = np.concatenate([Rm, t[:, None]], 1)
        depths.append(synth.render_depth(verts_w @ lw[:, :3].T + lw[:, 3], faces, K, 120, 160))
        lws.append(lw)
    fus = FusionDM_GPU(0.2, K, tsdf_res=R)
    tsdf, tsdfw = fus.compute_live_tsdf(depths, lws)
    assert tsdf.shape == (R, R, R) and tsdf.dtype == np.float32
    vox = ot.voxel_grid((R, R, R))
    ov, ow = np.full(R ** 3, 0.2), np.zeros(R ** 3)
    for dm, lw in zip(depths, lws):
        ov, ow, om, _ = ot.fuse_depth_rigid(ov, ow, vox, dm, lw, K, np.linalg.inv(K), 0.2, R, scale=scale, center=avg)
    assert ow.max() >= 2
    assert np.abs(tsdf.ravel() - ov).max() <= 1e-5 * 0.2 * 3
    assert np.array_equal(tsdfw.ravel(), ow.astype(np.float32))
    # numpy in -> numpy out fuseDepths keeps the caller's dtype and updates in place like the reference's nditer
    t64 = np.full((R, R, R), 0.2); w64 = np.zeros((R, R, R))
    rt, rw = fus.fuseDepths(depths[0], lws[0], t64, w64, scale=scale, center=avg)
    assert rt is t64 and rw is w64 and rt.dtype == np.float64
    o1, w1, _, _ = ot.fuse_depth_rigid(np.full(R ** 3, 0.2), np.zeros(R ** 3), vox, depths[0], lws[0], K, np.linalg.inv(K), 0.2, R, scale=scale, center=avg)
    assert np.abs(rt.ravel() - o1).max() <= 1e-5 * 0.2 and np.array_equal(rw.ravel(), w1)
    # device tensors in -> updated in place on the device, no host traffic
    td = torch.full((R, R, R), 0.2, device="cuda"); wd = torch.zeros((R, R, R), device="cuda")
    rt2, rw2 = fus.fuseDepths(torch.from_numpy(depths[0]).cuda(), lws[0], td, wd, scale=scale, center=avg)
    assert rt2 is td and np.abs(td.cpu().numpy().ravel() - o1).max() <= 1e-5 * 0.2


def test_fusion_updateTSDF_and_fuseFrame_match_oracle(small_scene):
    """Fusion.updateTSDF (a1, core/fusion.py:153-198) and Fusion.fuseFrame (a3) through the class surface, with the
    reference's initial float32 `_lw` and a 3-frame carry-over."""
    from dynamicfusion_body_b200 import Fusion, synth
    import scenes
    from oracle import tsdf as ot
    sc = small_scene
    R = sc.res
    vox, idx, tie = scenes.oracle_knn((R, R, R), sc.node_pos, sc.k)
    nw = np.full(sc.n_nodes, np.float32(sc.node_w))
    wv = synth.blend_warp(sc.vertices, sc.node_pos, sc.node_dq, nw, sc.vert_knn, lw=np.array([1, 0, 0, 0, 0, 0.1, 0, 0.]))
    live = synth.mesh_sdf_volume((R, R, R), wv, sc.warped_normals)
    tdist = float(live.max())
    fus = Fusion(tdist, knn=sc.k, use_cnn=False, write_warpfield=False)
    t0 = synth.mesh_sdf_volume((R, R, R), sc.vertices, sc.normals)
    fus.InitializeCanonicalSpace(tsdf=t0, K=sc.K, vertices=sc.vertices, normals=sc.normals, nodes=sc.nodes_as_reference_tuples())
    assert np.array_equal(fus.knn_indices()[~tie], idx[~tie])
    ov, ow = t0.ravel().astype(np.float64), np.zeros(R ** 3)
    for frame in range(3):
        fus.updateTSDF(live)                                   # a1 with _lw = [1,0,0,0,0,.1,0,0] float32 (core/fusion.py:57)
        ov, ow, om = ot.update_volume(ov, ow, live, vox, idx, sc.node_pos, sc.node_dq, nw, fus._lw, tdist)
    ok = ~tie
    assert np.abs(fus._tsdf.ravel() - ov)[ok].max() <= 1e-5 * tdist * 3
    assert (np.abs(fus._tsdfw.ravel() - ow) / np.maximum(1, ow))[ok].max() <= 1e-6
    # a3 through the class: switch the global dq to the camera and fuse the depth frame
    fus2 = Fusion(sc.tdist, knn=sc.k, use_cnn=False, write_warpfield=False)
    fus2.InitializeCanonicalSpace(tsdf_shape=(R, R, R), K=sc.K, vertices=sc.vertices, normals=sc.normals, nodes=sc.nodes_as_reference_tuples())
    fus2._lw = sc.lw
    m, f = fus2.fuseFrame(sc.depths, want_masks=True)
    ov, ow, om, ofr = ot.update_projective(np.full(R ** 3, sc.tdist), np.zeros(R ** 3), vox, idx, sc.node_pos, sc.node_dq, nw, sc.lw,
                                           sc.depths, sc.K, sc.Kinv, sc.tdist)
    assert np.array_equal((m.cpu().numpy() & 1).astype(bool)[ok], om[0][ok])
    assert np.abs(fus2._tsdf.ravel() - ov)[ok].max() <= 1e-5 * sc.tdist
    st = fus2.frame_stats()
    assert st["deferred"] == st["exact_processed"] and st["bricks"] > 0
