"""The reference's own driver loop (test.py:104-135) through the drop-in classes on the GPU:

    fus.setupCorrespondences(volume, method='clpts'); fus.solve(regularization_weight=0.5, method="clpts")
    fus.updateTSDF(); fus.update_graph()

over a short synthetic sequence (the body mesh of meshes/original.obj moved by a growing smooth deformation).  Surface
extraction -- skimage's marching cubes in the reference -- is supplied through the `surface_extractor` hook by a small
numpy zero-crossing extractor defined here (vertices = sign changes along grid edges, normals = TSDF gradient)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def edge_crossings(tsdf, step_size=1):
    """(verts, faces=None, normals, values=None): zero crossings of `tsdf` along the grid edges of a `step_size` lattice."""
    t = np.asarray(tsdf, dtype=np.float64)[::step_size, ::step_size, ::step_size]
    g = np.stack(np.gradient(t), -1)
    vs, ns = [], []
    for ax in range(3):
        lo = [slice(None)] * 3
        hi = [slice(None)] * 3
        lo[ax], hi[ax] = slice(None, -1), slice(1, None)
        lo, hi = tuple(lo), tuple(hi)
        cross = t[lo] * t[hi] < 0
        f = t[lo][cross] / (t[lo][cross] - t[hi][cross])
        p = np.argwhere(cross).astype(np.float64)
        p[:, ax] += f
        vs.append(p)
        ns.append(g[lo][cross] * (1 - f)[:, None] + g[hi][cross] * f[:, None])
    v = np.concatenate(vs) * step_size
    n = np.concatenate(ns)
    n /= np.maximum(np.linalg.norm(n, axis=1, keepdims=True), 1e-12)
    return v.astype(np.float32), None, n.astype(np.float32), None


def test_edge_crossings_helper():
    x, y, z = np.meshgrid(np.arange(20), np.arange(22), np.arange(24), indexing="ij")
    sdf = np.sqrt((x - 9.3) ** 2 + (y - 10.1) ** 2 + (z - 11.7) ** 2) - 6.0
    v, _, n, _ = edge_crossings(sdf)
    r = np.linalg.norm(v - np.array([9.3, 10.1, 11.7]), axis=1)
    assert len(v) > 300 and np.abs(r - 6.0).max() < 0.08
    assert (np.sum(n * (v - np.array([9.3, 10.1, 11.7])) / r[:, None], axis=1) > 0.97).all()


def test_reference_driver_loop():
    import torch
    from dynamicfusion_body_b200 import fusion, synth
    R = 48
    verts, normals, faces = synth.load_body_mesh()
    verts = (verts * (R - 1) / 64.0).astype(np.float32)
    tdist = 3.0

    def deform(p, a):
        """smooth bend + shift of magnitude a (voxels)"""
        q = p.astype(np.float64).copy()
        q[:, 0] += a * np.sin(p[:, 1] / R * np.pi)
        q[:, 2] += 0.5 * a * np.cos(p[:, 0] / R * np.pi)
        return q

    vol0 = synth.mesh_sdf_volume((R, R, R), verts, normals)
    fus = fusion.Fusion(float(vol0.max()), subsample_rate=1.5, knn=3, marching_cubes_step_size=2, verbose=False, use_cnn=False,
                        write_warpfield=False)                                            # test.py:110
    fus.surface_extractor = edge_crossings
    cv, _, cn, _ = edge_crossings(vol0, 2)
    fus.InitializeCanonicalSpace(tsdf=vol0, vertices=cv, normals=cn, radius=4.0)          # marching cubes + construct_graph
    fus._lw = np.array([1, 0, 0, 0, 0, 0, 0, 0], dtype=np.float64)
    n_nodes0 = len(fus._nodes)
    assert n_nodes0 > 20 and np.asarray(fus._neighbor_look_up).shape == (len(cv), 3)
    for it, a in enumerate((0.6, 1.2)):
        lv = deform(verts, a).astype(np.float32)
        live = synth.mesh_sdf_volume((R, R, R), lv, normals)
        fus.setupCorrespondences(live, method='clpts')                                    # test.py:122
        assert len(fus._correspondences) == len(fus._vertices) > 100
        v0 = fus._vertices.copy()
        before = np.linalg.norm(fus.warp(v0, locations=fus._neighbor_look_up, m_lw=fus._lw) - fus._correspondences, axis=1).mean()
        fus.solve(regularization_weight=0.5, method="clpts", gn_iterations=6)             # test.py:123
        assert fus.last_solve.cost <= fus.last_solve.cost0
        if len(fus._vertices) == len(v0):
            after = np.linalg.norm(fus.warp(fus._vertices, locations=fus._neighbor_look_up, m_lw=fus._lw) - fus._correspondences, axis=1).mean()
            assert after < before
        w_before = fus._tsdfw.sum()
        fus.updateTSDF()                                                                  # test.py:127  (a1, live volume kept by setupCorrespondences)
        assert fus._tsdfw.sum() > w_before and np.isfinite(fus._tsdf).all()
        fus.update_graph()                                                                # test.py:129
        assert len(fus._nodes) >= n_nodes0 and fus._curr_tsdf is None and fus._correspondences == []
        assert np.asarray(fus._neighbor_look_up).shape == (len(fus._vertices), 3)
        assert fus._node_vertex_idx.max() < len(fus._vertices) or len(fus._nodes) > n_nodes0
        n_nodes0 = len(fus._nodes)
    with pytest.raises(ValueError):
        fus.updateTSDF()                                                                  # live frame was released by update_graph (core/fusion.py:158)
    torch.cuda.synchronize()
