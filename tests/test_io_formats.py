"""On-disk formats: .dist SDF volumes, projection matrices, warp-field pickles."""
import os

import numpy as np
import pytest

from dynamicfusion_body_b200 import io
from oracle import refload


def _vol(rng):
    return rng.normal(size=(6, 5, 9)).astype(np.float32), rng.normal(size=(6, 5, 9, 3)).astype(np.float32)


def test_sdf_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    v, cp = _vol(rng)
    f = str(tmp_path / "0000.64.dist")
    io.save_sdf(f, v, b_min=(-1, -2, -3), b_max=(4, 5, 6), closest_points=cp)
    bmin, bmax, vol, cps = io.load_sdf(f, read_closest_points=True)
    assert np.array_equal(vol, v) and np.array_equal(cps, cp)
    assert np.array_equal(bmin, [-1, -2, -3]) and np.array_equal(bmax, [4, 5, 6])
    assert vol.shape == (6, 5, 9)
    # header layout of core/sdf.py:10-21: first two resolutions negated, x fastest in the payload
    raw = np.fromfile(f, dtype=np.int32, count=3)
    assert list(raw) == [-5, -4, 8]
    payload = np.fromfile(f, dtype=np.float32, offset=12 + 48, count=v.size)
    assert payload[1] == v[1, 0, 0] and payload[6] == v[0, 1, 0]
    with open(f, 'r+b') as fp:
        fp.truncate(100)
    with pytest.raises(ValueError):
        io.load_sdf(f)


@pytest.mark.reference
@pytest.mark.skipif(not refload.available(), reason="reference checkout not present")
def test_sdf_loader_matches_reference(tmp_path):
    refload.load()
    import core.sdf as rsdf
    rng = np.random.default_rng(1)
    v, cp = _vol(rng)
    f = str(tmp_path / "a.dist")
    io.save_sdf(f, v, closest_points=cp)
    r = rsdf.load_sdf(f, read_closest_points=True)
    m = io.load_sdf(f, read_closest_points=True)
    for a, b in zip(r, m):
        assert np.array_equal(a, b)


def test_proj_matrix_and_warp_field(tmp_path):
    p = tmp_path / "proj0.txt"
    P = np.arange(12, dtype=float).reshape(3, 4) * 0.5
    p.write_text("\n".join(" ".join(repr(float(x)) for x in row) for row in P) + "\n")
    assert np.array_equal(io.read_proj_matrix(str(p)), P)
    nodes = [(3, np.zeros(3, np.float32), np.array([1, 0, 0, 0, 0, .01, .01, 0], np.float32), 2.5)]
    io.write_warp_field(nodes, str(tmp_path), "test", 7)
    back = io.read_warp_field(os.path.join(str(tmp_path), "test__7.p"))
    assert back[0][0] == 3 and np.array_equal(back[0][2], nodes[0][2]) and back[0][3] == 2.5


@pytest.mark.skipif(not refload.available(), reason="reference checkout not present")
def test_obj_writer_matches_reference(tmp_path):
    """io.write_obj against the unmodified Fusion.write_canonical_mesh / FusionDM.write_canonical_mesh (core/fusion.py:577-586,
    core/fusion_dm.py:339-354) with skimage's marching cubes stubbed to return a fixed mesh."""
    import sys
    util, Fusion, FusionDM = refload.load()
    rng = np.random.default_rng(2)
    verts = (rng.random((30, 3)) * 20).astype(np.float32); normals = rng.normal(size=(30, 3)).astype(np.float32)
    faces = rng.integers(0, 30, size=(40, 3)).astype(np.int32)
    sys.modules["skimage.measure"].marching_cubes_lewiner = lambda *a, **kw: (verts, faces, normals, None)
    try:
        f = refload.make_fusion([(0, np.zeros(3, np.float32), np.array([1, 0, 0, 0, 0, 0, 0, 0], np.float32), 1.0)] * 5, np.zeros((2, 2, 2)),
                                np.zeros((2, 2, 2)), 1.0, 4, None)
        f.write_canonical_mesh(str(tmp_path), "ref.obj")
        io.write_obj(str(tmp_path / "mine.obj"), verts, normals, faces)
        assert (tmp_path / "ref.obj").read_text() == (tmp_path / "mine.obj").read_text()
        fdm = FusionDM(0.5, np.eye(3), tsdf_res=4)
        fdm._IND = np.array([[0.5, 0, 0, 1.0], [0, 0.5, 0, -2.0], [0, 0, 0.5, 0.25], [0, 0, 0, 1]])
        fdm.write_canonical_mesh(str(tmp_path), "refdm.obj")
        rot, trans = fdm._IND[:3, :3], fdm._IND[:3, 3]
        io.write_obj(str(tmp_path / "minedm.obj"), verts.astype(np.float64) @ rot.T + trans, normals.astype(np.float64) @ rot.T, faces, face_normals=True)
        assert (tmp_path / "refdm.obj").read_text() == (tmp_path / "minedm.obj").read_text()
    finally:
        del sys.modules["skimage.measure"].marching_cubes_lewiner
