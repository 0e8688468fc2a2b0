"""GPU parity tests (run with -m gpu) of SURVEY 8f ranks 1-2 -- closest-point correspondences and deformation-graph
maintenance (csrc/graph.cu behind Fusion / FusionDM.setupCorrespondences, update_graph, construct_graph) -- against
the golden vectors produced by executing the unmodified reference (tests/golden/make_golden_graph.py) and against
oracle/graph.py on randomised inputs.  Bars: neighbour / correspondence / sample indices bit-exact (exact distance ties
excluded), correspondences bit-identical float32 points, blended node transforms to float32 storage precision."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GG = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_graph_vectors.npz"))


def _fusion(cls_name="Fusion"):
    from dynamicfusion_body_b200 import fusion
    k, radius = int(GG["k"]), float(GG["radius"])
    if cls_name == "Fusion":
        f = fusion.Fusion(1.0, knn=k, use_cnn=False, write_warpfield=False)
        N = len(GG["node_pos"])
        f._vertices, f._normals, f._radius = GG["verts"].copy(), GG["norms"].copy(), radius
        f._nodes = [(int(GG["node_idx"][i]), GG["node_pos"][i], GG["node_dq"][i], 2 * radius) for i in range(N)]
        f._neighbor_look_up = f._lookup(f._vertices, k).astype(np.int64)
    else:
        f = fusion.FusionDM(0.5, np.eye(3), tsdf_res=8, knn=k, write_warpfield=False)
        f._vertices, f._normals = GG["verts"].copy(), GG["norms"].copy()
    f._lw = GG["lw"].copy()
    return f


@pytest.mark.parametrize("k", [1, 4, 8])
def test_point_grid_knn_matches_bruteforce(k):
    from dynamicfusion_body_b200 import engine
    from oracle import graph as og
    rng = np.random.default_rng(k)
    pts = (rng.random((6000, 3)) * np.array([40, 25, 10]) + 3).astype(np.float32)
    pts[100:400] = pts[100]                                         # a heavy cell of coincident points (exact ties)
    q = np.concatenate([rng.random((1500, 3)) * np.array([40, 25, 10]) + 3,          # inside the box
                        rng.random((300, 3)) * 120 - 40,                             # partly far outside
                        pts[:200].astype(np.float64)])                               # exactly on data points
    for cell in (None, 0.7, 9.0):
        grid = engine.PointGrid(pts, cell=cell)
        idx, d2 = grid.knn(q, k, want_d2=True)
        idx, d2 = idx.cpu().numpy(), d2.cpu().numpy()
        oi, od2 = og.knn_points(pts, q, k)
        assert np.array_equal(d2, od2)                              # distances agree bit for bit, ties included
        assert np.array_equal(idx, oi)                              # ties resolved by lower id on both sides


def test_point_grid_small_sets():
    from dynamicfusion_body_b200 import engine
    pts = np.array([[1, 2, 3], [1, 2, 4]], np.float32)
    idx = engine.PointGrid(pts).knn(np.array([[1.0, 2.0, 3.9]]), 4).cpu().numpy()
    assert idx.tolist() == [[1, 0, -1, -1]]
    idx = engine.PointGrid(np.zeros((0, 3), np.float32)).knn(np.zeros((3, 3)), 2).cpu().numpy()
    assert (idx == -1).all()


def test_vertex_node_table_golden():
    f = _fusion()
    assert np.array_equal(np.asarray(f._neighbor_look_up), GG["vknn"])


def test_setupCorrespondences_golden():
    f = _fusion()
    f.setupCorrespondences(None, method='clpts', prune_result=False, live_vertices=GG["lverts"])
    assert f._correspondences.dtype == np.float32
    assert np.array_equal(f._correspondences, GG["corr_fusion"])


def test_setupCorrespondences_prune_and_errors():
    from oracle import dq as odq
    from oracle import graph as og
    f = _fusion()
    k, radius = int(GG["k"]), float(GG["radius"])
    vknn = GG["vknn"]
    wv, wn = odq.warp(GG["verts"], GG["node_pos"][vknn], GG["node_dq"][vknn], np.full(vknn.shape, 2 * radius), lw=GG["lw"], normal=GG["norms"])
    nn, _ = og.knn_points(GG["lverts"], wv, k)
    best, cost = og.corr_select(wv, wn, GG["lverts"], nn)
    keep = cost <= 0.2
    assert 0 < keep.sum() < len(keep)
    f.setupCorrespondences(None, method='clpts', live_vertices=GG["lverts"])          # prune_result=True, tolerance=0.2
    assert np.array_equal(f._vertices, GG["verts"][keep]) and np.array_equal(f._normals, GG["norms"][keep])
    assert np.array_equal(f._correspondences, GG["lverts"][best[keep]])
    assert np.array_equal(np.asarray(f._neighbor_look_up), vknn[keep])
    link, _ = og.knn_points(GG["verts"][keep], GG["node_pos"], 1)
    assert np.array_equal(f._node_vertex_idx, link[:, 0])
    assert len(f._correspondences) == len(f._vertices)                                  # solve()'s precondition (core/fusion.py:337)
    f.surface_extractor = None
    with pytest.raises(NotImplementedError):
        f.setupCorrespondences(np.zeros((4, 4, 4)))                                     # no surface extractor configured
    with pytest.raises(ValueError):
        f.setupCorrespondences(None, live_vertices=GG["lverts"][:2])
    # the hook takes the place of skimage's marching cubes
    f.surface_extractor = lambda tsdf, step: (GG["lverts"], None, None, None)
    f.setupCorrespondences(np.zeros((4, 4, 4)), prune_result=False)
    assert len(f._correspondences) == keep.sum()


def test_fusiondm_setupCorrespondences_golden():
    f = _fusion("FusionDM")
    f.setupCorrespondences(None, tolerance=float(GG["dm_tolerance"]), live_vertices=GG["lverts"])
    assert np.array_equal(np.array(f._corridx), GG["dm_corridx"])
    assert np.array_equal(np.array(f._correspondences), GG["dm_corr"])


@pytest.mark.parametrize("radius", [0.05, 1.3, 2.5, 6.0, 500.0])
def test_uniform_sample_matches_sequential(radius):
    from dynamicfusion_body_b200 import engine
    from oracle import graph as og
    pts = GG["lverts"][:1200]
    v, i = engine.uniform_sample(pts, radius)
    ov, oi = og.uniform_sample(pts, radius)
    assert np.array_equal(i, oi) and np.array_equal(v, ov)
    if radius == 2.5:
        assert np.array_equal(i, GG["us_idx"])
    # a scan-line ordered point set (long dependency chains between rounds)
    g = np.stack(np.meshgrid(np.arange(40), np.arange(30), np.arange(3), indexing="ij"), -1).reshape(-1, 3).astype(np.float32) * 0.5
    v, i = engine.uniform_sample(g, radius)
    ov, oi = og.uniform_sample(g, radius)
    assert np.array_equal(i, oi)


def test_construct_graph_golden():
    from dynamicfusion_body_b200 import fusion
    f = fusion.Fusion(1.0, knn=int(GG["k"]), use_cnn=False, write_warpfield=False)
    f.InitializeCanonicalSpace(tsdf_shape=(8, 8, 8), vertices=GG["verts"], normals=GG["norms"], radius=float(GG["radius"]))
    nodes = f._nodes
    assert np.array_equal(np.array([n[0] for n in nodes]), GG["node_idx"])
    assert np.array_equal(np.array([n[1] for n in nodes]), GG["node_pos"])
    assert np.array_equal(nodes[0][2], np.array([1, 0, 0, 0, 0, 0.01, 0.01, 0], np.float32))          # Q5
    assert nodes[0][3] == 2 * float(GG["radius"])
    assert np.array_equal(np.asarray(f._neighbor_look_up), GG["vknn"])


def test_update_graph_golden():
    f = _fusion()
    N = len(GG["node_pos"])
    f._correspondences = [1]
    f.update_graph(vertices=GG["ug_verts"])
    nodes = f._nodes
    assert len(nodes) == len(GG["ug_node_pos"]) > N
    assert np.array_equal(np.array([n[0] for n in nodes]), GG["ug_node_vidx"])
    assert np.array_equal(np.array([n[1] for n in nodes]), GG["ug_node_pos"])
    dq = np.array([n[2] for n in nodes])
    assert np.array_equal(dq[:N], GG["node_dq"])
    assert np.abs(dq[N:] - GG["ug_node_dq"][N:]).max() <= 6e-8 * np.abs(GG["ug_node_dq"][N:]).max()  # float32 storage of a float64 blend
    assert np.array_equal(np.asarray(f._neighbor_look_up), GG["ug_lookup"])
    assert f._curr_tsdf is None and f._correspondences == []
    f.surface_extractor = None
    with pytest.raises(NotImplementedError):
        f.update_graph()                                                                             # no surface extractor


def test_fusiondm_solve_recovers_rigid_motion():
    """FusionDM.solve (core/fusion_dm.py:262-281): canonical surface = a subset of the live surface moved by a known rigid
    transform; three rounds of closest-point correspondences + rigid least squares must bring it back."""
    from dynamicfusion_body_b200 import fusion, synth
    from oracle import dq as odq
    mesh = np.load(os.path.join(os.path.dirname(__file__), "golden", "body_mesh.npz"))
    live = mesh["vertices"]
    ang = 0.03
    q = np.array([np.cos(ang / 2), 0, np.sin(ang / 2), 0])
    t = np.array([0.25, -0.15, 0.2])
    true_lw = np.concatenate([q, 0.5 * odq.quaternion_multiply(np.array([0, *t]), q)])
    inv = np.concatenate([q * [1, -1, -1, -1], -0.5 * odq.quaternion_multiply(q * [1, -1, -1, -1], np.array([0, *t]))])
    canon = odq.dqb_warp(inv, live[::9]).astype(np.float32)                      # canonical = T^-1 (live)
    cn = odq.dqb_warp_normal(inv, mesh["normals"][::9]).astype(np.float32)
    f = fusion.FusionDM(0.5, np.eye(3), tsdf_res=8, knn=4, write_warpfield=False)
    f._vertices, f._normals = canon, cn
    r0 = np.abs(odq.dqb_warp(np.asarray(f._lw, np.float64), canon) - live[::9]).max()
    f.solve(None, live_vertices=live)
    r1 = np.abs(odq.dqb_warp(f._lw, canon) - live[::9]).max()
    assert f._lw.dtype == np.float64 and r0 > 0.2 and r1 < 0.02
    res = f.computef_lw(f._lw)
    assert res.shape == (len(f._corridx),) and np.abs(res).max() < 0.05


def test_fusion_solve_clpts_flow():
    """Fusion.solve(method='clpts') re-runs the correspondence search between solver rounds (core/fusion.py:363-371)."""
    f = _fusion()
    f.surface_extractor = lambda tsdf, step: (GG["lverts"], None, None, None)
    f._curr_tsdf = np.zeros((4, 4, 4), np.float32)
    f.setupCorrespondences(f._curr_tsdf, method='clpts')
    n0 = len(f._vertices)
    f.solve(method='clpts', gn_iterations=3)
    assert len(f._correspondences) == len(f._vertices) <= n0
    assert np.isfinite(f.last_solve.cost) and f.last_solve.cost <= f.last_solve.cost0


def test_incremental_graph_revision_equals_full_rebuild():
    """update_graph only appends nodes (core/fusion.py:216-229): DeviceWarpField.append_nodes brings the voxel kNN table and the brick /
    region candidate sets up to date by rebuilding only the 8^3 bricks a new node can reach.  The table must equal a full rebuild bit
    for bit, and a fused frame on the incrementally updated field must equal one on a freshly built field."""
    import torch
    from dynamicfusion_body_b200 import engine, synth
    sc = synth.make_scene(res=96, k=4, n_nodes=400, seed=2, rows=120, cols=160, background=True)
    R = sc.res
    n0 = sc.n_nodes - 9                                  # the last 9 nodes arrive as a graph revision (+2 %)
    w = np.float32(sc.node_w)
    wf = engine.DeviceWarpField(4)
    wf.set_nodes(sc.node_pos[:n0], sc.node_dq[:n0], w)
    depth = torch.from_numpy(sc.depths).cuda()
    vol = engine.DeviceVolume((R, R, R), fill=sc.tdist)
    engine.update_projective(vol, wf, sc.lw, depth, sc.K, sc.Kinv, sc.extrinsics, sc.tdist)      # builds + uses the tables of the old graph
    wf.append_nodes(sc.node_pos[n0:], sc.node_dq[n0:], w)
    dirty = wf.last_dirty.cpu().numpy()
    assert 0 < dirty.mean() < 0.6                        # only part of the volume is rebuilt
    ref = engine.DeviceWarpField(4)
    ref.set_nodes(sc.node_pos, sc.node_dq, w)
    assert torch.equal(wf.knn_table((R, R, R), 0, R), ref.knn_table((R, R, R), 0, R))
    b_inc, b_ref = wf.brick_nodes((R, R, R), 0, R), ref.brick_nodes((R, R, R), 0, R)
    assert torch.equal(b_inc[1], b_ref[1]) and torch.equal(b_inc[4], b_ref[4])                  # candidate counts per brick / region
    for cnt, a, b in ((b_inc[1], b_inc[0], b_ref[0]), (b_inc[4], b_inc[3], b_ref[3])):          # same node SETS (list order depends on atomics)
        a, b, c = a.cpu().numpy().view(np.uint16).astype(np.int64), b.cpu().numpy().view(np.uint16).astype(np.int64), cnt.cpu().numpy()
        width = a.shape[1]
        valid = (np.arange(width)[None, :] < np.minimum(c, width)[:, None]) & (c[:, None] <= width)
        assert np.array_equal(np.sort(np.where(valid, a, -1), axis=1), np.sort(np.where(valid, b, -1), axis=1))
    vol_ref = engine.DeviceVolume((R, R, R), tsdf=vol.tsdf.clone(), weight=vol.weight.clone())
    m1 = engine.update_projective(vol, wf, sc.lw, depth, sc.K, sc.Kinv, sc.extrinsics, sc.tdist, want_masks=True)
    m2 = engine.update_projective(vol_ref, ref, sc.lw, depth, sc.K, sc.Kinv, sc.extrinsics, sc.tdist, want_masks=True)
    assert torch.equal(vol.tsdf, vol_ref.tsdf) and torch.equal(vol.weight, vol_ref.weight) and torch.equal(m1[0], m2[0]) and torch.equal(m1[1], m2[1])
