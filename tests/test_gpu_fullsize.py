"""Full-size GPU checks (BASELINE configs): the size-independent properties plus oracle parity on voxel samples.

 * config 0: 128^3 rigid fusion (identity warp) with the reference's intrinsics K (test.py:141) -- whole volume vs oracle
 * config 1/4: 256^3 / 512^3 warped projective update -- hybrid (brick culling + fp32 tier) == all-exact bit for bit,
   oracle parity on a 200k-voxel random sample, slab concatenation == full volume, update idempotence of SKIP voxels
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _engine():
    import torch
    from dynamicfusion_body_b200 import engine
    return torch, engine


def test_config0_rigid_128_reference_intrinsics():
    torch, engine = _engine()
    from dynamicfusion_body_b200 import synth
    from oracle import tsdf as ot
    R = 128
    K = np.array([2000, 0, 800, 0, 2000, 600, 0, 0, 1], dtype=float).reshape(3, 3)     # test.py:141
    Kinv = np.linalg.inv(K)
    v64, nrm, faces = synth.load_body_mesh()
    std, avg = 1.3, np.array([-0.03, -0.43, -5.6])                                      # core/fusion_dm.py:106-107
    scale = 12 * std / R
    verts_w = scale * (v64.astype(np.float64) * (R - 1) / 64.0 - R / 2) + avg
    lw34 = np.concatenate([np.eye(3), np.array([[0.0], [0.4], [0.0]])], 1)              # identity warp, camera at the origin
    lw34[:, 3] = -avg * 0 + np.array([0.03, 0.43, 0.0])
    cam = verts_w @ lw34[:, :3].T + lw34[:, 3]
    cam[:, 2] *= -1                                                                      # look down -z of the data: flip into +z
    lw34 = np.diag([1.0, 1.0, -1.0]) @ lw34
    dm = synth.render_depth(verts_w @ lw34[:, :3].T + lw34[:, 3], faces, K, 1200, 1600)
    assert (dm != 0).mean() > 0.05
    tdist = 0.2                                                                          # test.py:159
    vox = ot.voxel_grid((R, R, R))
    t0 = np.full(R ** 3, tdist, np.float32); w0 = np.zeros(R ** 3, np.float32)
    ov, ow, om, ofr = ot.fuse_depth_rigid(t0.astype(np.float64), w0.astype(np.float64), vox, dm, lw34, K, Kinv, tdist, R, scale=scale, center=avg)
    assert om.mean() > 0.01
    vol = engine.DeviceVolume((R, R, R), tsdf=t0, weight=w0)
    mask, frus = engine.fuse_depth_rigid(vol, R, torch.from_numpy(dm).cuda(), lw34, K, Kinv, scale, avg, tdist, want_masks=True)
    assert np.array_equal(mask.cpu().numpy().astype(bool), om) and np.array_equal(frus.cpu().numpy().astype(bool), ofr)
    assert np.abs(vol.tsdf.cpu().numpy().ravel() - ov).max() <= 1e-5 * tdist
    assert np.array_equal(vol.weight.cpu().numpy().ravel(), ow.astype(np.float32))
    print("config 0: updated", om.mean(), "frustum", ofr.mean(), vol.workspace.stats())


@pytest.mark.parametrize("R,N,k,views", [(256, 1000, 4, 1), (512, 4000, 4, 1), (256, 1000, 8, 8)])
def test_full_size_projective(R, N, k, views):
    torch, engine = _engine()
    from dynamicfusion_body_b200 import synth
    from oracle import driver, tsdf as ot
    sc = synth.make_scene(res=R, k=k, n_nodes=N, seed=0, background=True, n_views=views)
    wf = engine.DeviceWarpField(k)
    wf.set_nodes(sc.node_pos, sc.node_dq, np.float32(sc.node_w))
    depths = torch.from_numpy(sc.depths).cuda()
    rng = np.random.default_rng(0)
    t0 = torch.from_numpy((rng.normal(size=R ** 3) * 0.3 * sc.tdist).clip(-sc.tdist, sc.tdist).astype(np.float32)).cuda()
    w0 = torch.from_numpy(rng.integers(0, 110, size=R ** 3).astype(np.float32)).cuda()
    res = {}
    for name, mode in (("hybrid", 0), ("exact", 1)):
        vol = engine.DeviceVolume((R, R, R), tsdf=t0.clone(), weight=w0.clone())
        m, f = engine.update_projective(vol, wf, sc.lw, depths, sc.K, sc.Kinv, sc.extrinsics, sc.tdist, mode=mode, want_masks=True)
        res[name] = (vol.tsdf, vol.weight, m, f)
        if mode == 0:
            print("R=%d views=%d" % (R, views), vol.workspace.stats())
    # the classification tiers never change a decision, and a clamped update is the same float32 expression in both tiers
    assert torch.equal(res["hybrid"][2], res["exact"][2]) and torch.equal(res["hybrid"][3], res["exact"][3])
    assert torch.equal(res["hybrid"][1], res["exact"][1])
    assert torch.equal(res["hybrid"][0], res["exact"][0])
    # untouched voxels are bit-identical to the input (idempotence of SKIP)
    untouched = res["hybrid"][2] == 0
    assert torch.equal(res["hybrid"][0][untouched.view(R, R, R)], t0.view(R, R, R)[untouched.view(R, R, R)])
    # oracle parity on a voxel sample
    lin, vox = driver.sample_voxels((R, R, R), 200_000, seed=1)
    sd = driver.scene_dict(sc)
    ov, ow, om, ofr = driver.projective_on_sample(sd, vox, t0.cpu().numpy()[lin].astype(np.float64), w0.cpu().numpy()[lin].astype(np.float64))
    gm = res["hybrid"][2].cpu().numpy()[lin]
    for v in range(views):
        assert np.array_equal(((gm >> v) & 1).astype(bool), om[v])
    assert np.abs(res["hybrid"][0].cpu().numpy().ravel()[lin] - ov).max() <= 1e-5 * sc.tdist
    assert (np.abs(res["hybrid"][1].cpu().numpy().ravel()[lin] - ow) / np.maximum(1, ow)).max() <= 1e-6
    # slab sharding: two slabs == full volume
    h = R // 2 + 3
    parts = []
    for x0, x1 in ((0, h), (h, R)):
        s = engine.DeviceVolume((R, R, R), x0, x1, tsdf=t0.view(R, R, R)[x0:x1].clone(), weight=w0.view(R, R, R)[x0:x1].clone())
        engine.update_projective(s, wf, sc.lw, depths, sc.K, sc.Kinv, sc.extrinsics, sc.tdist)
        parts.append((s.tsdf, s.weight))
    # bit for bit, whichever tier the brick grid of a slab routes a voxel through
    assert torch.equal(torch.cat([p[1] for p in parts]), res["hybrid"][1])
    assert torch.equal(torch.cat([p[0] for p in parts]), res["hybrid"][0])
