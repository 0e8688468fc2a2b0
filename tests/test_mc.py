"""Surface extraction.  The reference delegates to skimage.measure.marching_cubes_lewiner (absent here, no golden
mesh in the reference) -> parity unpinned; the mesh is defined by oracle/mc.py.  CPU: the oracle's own properties, the generated
table against the oracle's first-principles derivation, and the host build of csrc/dfb_mc.h (same functions the kernels run)
against the oracle.  GPU: the C-ABI path against the oracle."""
import os
import re
from collections import Counter

import numpy as np
import pytest

from oracle import mc as omc

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sphere(res=24, c=(11.3, 12.1, 11.7), r=8.2, shape=None):
    shape = shape or (res, res, res)
    x, y, z = np.meshgrid(*[np.arange(n) for n in shape], indexing="ij")
    return (np.sqrt((x - c[0]) ** 2 + (y - c[1]) ** 2 + (z - c[2]) ** 2) - r).astype(np.float32)


def mesh_topology(verts, faces, weld=False):
    """(#directed edges used more than once, #directed edges without their opposite, Euler characteristic).  weld: identify
    vertices by position first -- edge vertices snapped onto one grid sample (a sample value exactly at the level) coincide."""
    if weld:
        _, inv = np.unique(verts, axis=0, return_inverse=True)
        faces = inv.reshape(-1)[faces]
    E = Counter()
    for a, b, c in faces:
        for p, q in ((a, b), (b, c), (c, a)):
            E[(int(p), int(q))] += 1
    dup = sum(1 for n in E.values() if n != 1)
    open_ = sum(1 for (p, q) in E if (q, p) not in E)
    used = len(set(int(i) for i in faces.ravel()))
    return dup, open_, used - len(E) // 2 + len(faces)


def volumes():
    rng = np.random.default_rng(5)
    two = np.minimum(sphere(20, (6.2, 7.1, 6.6), 4.3), sphere(20, (13.4, 12.2, 12.9), 4.1))
    snapped = sphere(16, (8, 8, 8), 5.0)                                   # integer centre/radius: samples exactly at the level
    noise = rng.normal(size=(9, 7, 37)).astype(np.float32)                 # every ambiguous configuration, ragged chunk
    clipped = np.clip(sphere(40, (19.5, 20.2, 18.8), 13.0, shape=(40, 36, 70)), -3, 3).astype(np.float32)
    return {"sphere": (sphere(), 1, 0.0), "sphere_step2": (sphere(33, (16.2, 15.7, 16.4), 11.0), 2, 0.0),
            "sphere_step3_auto": (sphere(31, (15.2, 15.7, 14.4), 9.0), 3, None), "two": (two, 1, 0.0),
            "snapped": (snapped, 1, 0.0), "noise": (noise, 1, 0.1), "noise_step2": (noise, 2, None),
            "clipped_auto": (clipped, 1, None), "flat": (np.ones((5, 6, 7), np.float32), 1, None),
            "tiny": (np.array([[[0, 1], [1, 1]], [[1, 1], [1, 2]]], np.float32) - 0.5, 1, 0.0)}


def assert_same_mesh(got, want, name):
    v, f, n, val = got
    ov, of, on, oval = want
    assert v.shape == ov.shape and f.shape == of.shape, name
    assert np.array_equal(f, of), name                                     # index work: bit-exact
    assert np.array_equal(v, ov), name                                     # float32 statement with explicit roundings: bit-exact
    assert np.array_equal(val, oval), name
    assert np.allclose(n, on, rtol=0, atol=1e-6), name                     # tolerance for the unit normals: 1e-6 absolute


def test_generated_table_is_current_and_matches_oracle():
    text = open(os.path.join(_ROOT, "dynamicfusion_body_b200", "csrc", "dfb_mc_table.h")).read()
    ntri = [int(x) for x in re.search(r"DFB_MC_NTRI_INIT \{(.*?)\}", text, re.S).group(1).replace("\\", "").split(",") if x.strip()]
    tri = [int(x) for x in re.search(r"DFB_MC_TRI_INIT \{(.*?)\}", text, re.S).group(1).replace("\\", "").split(",") if x.strip()]
    width = len(tri) // 256
    table = omc.case_table()
    assert len(ntri) == 256 and width == 15
    for c in range(256):
        assert ntri[c] == len(table[c])
        flat = [e for t in table[c] for e in t]
        assert tri[c * width:c * width + len(flat)] == flat and all(e == -1 for e in tri[c * width + len(flat):(c + 1) * width])


def test_oracle_mesh_properties():
    v, f, n, val = omc.marching_cubes(sphere(), 1, level=0.0)
    assert mesh_topology(v, f) == (0, 0, 2)                                # closed, consistently oriented, genus 0
    assert np.abs(np.linalg.norm(v - [11.3, 12.1, 11.7], axis=1) - 8.2).max() < 0.03
    fn = np.cross(v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 0]])
    assert ((fn * n[f[:, 0]]).sum(1) > 0).all()                            # winding agrees with the vertex normals
    assert abs(0.5 * np.linalg.norm(fn, axis=1).sum() / (4 * np.pi * 8.2 ** 2) - 1) < 0.01
    radial = (v - [11.3, 12.1, 11.7]) / 8.2
    assert ((radial * n).sum(1) > 0.99).all()                              # +gradient of a distance field = outward
    vols = volumes()
    v, f, _, _ = omc.marching_cubes(*vols["two"])
    assert mesh_topology(v, f, weld=True) == (0, 0, 4)                     # two spheres (two grid samples lie exactly on one)
    v, f, _, _ = omc.marching_cubes(*vols["snapped"])
    dup, open_, chi = mesh_topology(v, f, weld=True)
    assert dup == 0 and open_ == 0 and chi == 2 and len(f) > 0                         # dropping snapped triangles keeps the surface closed
    fn = np.cross(v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 0]])
    keys = [tuple(map(tuple, v[t])) for t in f]
    assert all(len(set(k)) == 3 for k in keys)                             # no triangle with coincident vertices
    # noise: every ambiguous configuration.  Fan diagonals of two cells can coincide on a shared ambiguous face (an edge used twice in
    # each direction), so the check is the cycle property: away from the volume border every directed edge is matched by its opposite
    v, f, _, _ = omc.marching_cubes(*vols["noise"])
    E = Counter((int(a), int(b)) for t in f for a, b in ((t[0], t[1]), (t[1], t[2]), (t[2], t[0])))
    inside = lambda i: all(0 < v[i][a] < n - 1 for a, n in enumerate((9, 7, 37)))
    assert len(E) > 3000 and all(E[(q, p)] == c for (p, q), c in E.items() if inside(p) or inside(q))
    v, f, _, _ = omc.marching_cubes(*vols["flat"])
    assert len(v) == 0 and f.shape == (0, 3)
    v, f, _, _ = omc.marching_cubes(*vols["tiny"])
    assert len(v) == 3 and len(f) == 1
    assert omc.default_level(vols["clipped_auto"][0]) == 0.0


def test_host_build_of_kernel_logic_matches_oracle():
    import hostshim_api as hs
    for name, (vol, step, level) in volumes().items():
        assert_same_mesh(hs.marching_cubes(vol, step, level), omc.marching_cubes(vol, step, level), name)


@pytest.mark.gpu
def test_device_marching_cubes_matches_oracle():
    import torch
    from dynamicfusion_body_b200 import engine
    for name, (vol, step, level) in volumes().items():
        got = engine.marching_cubes(torch.from_numpy(vol).cuda(), step, level)
        assert_same_mesh(got, omc.marching_cubes(vol, step, level), name)
    with pytest.raises(ValueError):
        engine.marching_cubes(torch.zeros(4, 4, device="cuda"), 1)
    with pytest.raises(ValueError):
        engine.marching_cubes(torch.zeros(4, 4, 4, device="cuda"), 0)


@pytest.mark.gpu
def test_device_marching_cubes_full_size_properties():
    """256^3 (BASELINE config 1 grid): closed genus-0 surface of the right area; step 2 halves the sampling."""
    import torch
    from dynamicfusion_body_b200 import engine
    R, c, r = 256, (127.3, 128.1, 126.7), 90.2
    ax = torch.arange(R, device="cuda", dtype=torch.float32)
    vol = torch.sqrt((ax[:, None, None] - c[0]) ** 2 + (ax[None, :, None] - c[1]) ** 2 + (ax[None, None, :] - c[2]) ** 2) - r
    for step, tol in ((1, 2e-3), (2, 4e-3)):
        v, f, n, _ = engine.marching_cubes(vol, step, 0.0)
        assert np.abs(np.linalg.norm(v - np.array(c), axis=1) - r).max() < 0.01 * step * step
        fn = np.cross(v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 0]])
        assert abs(0.5 * np.linalg.norm(fn, axis=1).sum() / (4 * np.pi * r * r) - 1) < tol
        assert ((fn * n[f[:, 0]]).sum(1) > 0).all()
        e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]]).astype(np.int64)
        fwd = np.unique(e[:, 0] * len(v) + e[:, 1]); bwd = np.unique(e[:, 1] * len(v) + e[:, 0])
        assert len(fwd) == len(e) and np.array_equal(fwd, bwd)             # every directed edge once, with its opposite
        assert len(v) - len(e) // 2 + len(f) == 2


# ---- x-slabs (SURVEY 8e x 8f rank 3): the slabs' meshes concatenate to the single-volume mesh bit for bit --------------------
def _slab_cases():
    rng = np.random.default_rng(11)
    noise = rng.normal(size=(23, 6, 35)).astype(np.float32)
    snapped = sphere(20, (10, 10, 10), 6.0)                                # samples exactly at the level: degenerate-triangle decisions
    return [("sphere_w2", sphere(), 1, 0.0, 2), ("sphere_w3_step2", sphere(33, (16.2, 15.7, 16.4), 11.0), 2, 0.0, 3),
            ("noise_w4", noise, 1, 0.1, 4), ("noise_w3_step3", noise, 3, 0.05, 3), ("snapped_w5", snapped, 1, 0.0, 5)]


def _slabwise(extract, vol, step, level, world):
    """extract_surface_slab's single-process composition: every slab with the halo planes its neighbours would send."""
    import torch
    from dynamicfusion_body_b200 import dist as ddist
    rx = vol.shape[0]
    parts, offset = [], 0
    for r, (x0, x1) in enumerate(ddist.slab_partition(rx, world)):
        a, b, _ = ddist.slab_sample_planes(x0, x1, rx, step)
        prev = torch.from_numpy(vol[a - step]) if r > 0 else None
        nxt = torch.from_numpy(np.stack([vol[b + step], vol[b + 2 * step]])) if r < world - 1 else None
        sub, lo, a, b = ddist.slab_halo_volume(torch.from_numpy(vol[x0:x1]), x0, x1, rx, step, prev, nxt)
        v, f, n, val = ddist.cut_owned_mesh(extract(sub, step, level, x_origin=lo // step, plane_offsets=True), lo, a, b, step)
        parts.append((v, (f + offset).astype(np.int32), n, val))
        offset += len(v)
    return tuple(np.concatenate([p[i] for p in parts]) for i in range(4))


def _assert_identical(got, want, name):
    for g, w in zip(got, want):
        assert g.shape == w.shape and np.array_equal(g, w), name


def test_slab_meshes_concatenate_to_the_full_mesh_host_build():
    import hostshim_api as hs
    for name, vol, step, level, world in _slab_cases():
        full = hs.marching_cubes(vol, step, level)
        assert len(full[1]) > 50, name
        got = _slabwise(lambda sub, s, lv, **kw: hs.marching_cubes(sub.numpy(), s, lv, **kw), vol, step, level, world)
        _assert_identical(got, full, name)


def test_x_origin_matches_oracle_and_moves_the_degenerate_decisions():
    import hostshim_api as hs
    rng = np.random.default_rng(3)
    vol = rng.normal(size=(6, 7, 9)).astype(np.float32)
    vol[2, 3, 4] = np.float32(0.1) + np.float32(5e-6)                      # x edge: t = 1e-5, distinct from the sample at x index 2,
    vol[3, 3, 4] = np.float32(-0.4)                                        # rounds onto it at x index 2 + 4096;
    vol[2, 4, 4] = np.float32(-100)                                        # y edge: t = 5e-8, on the sample either way
    for xo in (0, 5, 4096):
        assert_same_mesh(hs.marching_cubes(vol, 1, 0.1, x_origin=xo), omc.marching_cubes(vol, 1, 0.1, x_origin=xo), "xo%d" % xo)
    v0, f0 = hs.marching_cubes(vol, 1, 0.1)[:2]
    v1, f1 = hs.marching_cubes(vol, 1, 0.1, x_origin=4096)[:2]
    assert np.array_equal(v1[:, 1:], v0[:, 1:]) and np.abs(v1[:, 0] - 4096 - v0[:, 0]).max() < 1e-3
    assert len(f1) == len(f0) - 4                                          # triangles that only degenerate at the large origin
    with pytest.raises(ValueError):
        from dynamicfusion_body_b200 import dist as ddist
        ddist.slab_sample_planes(4, 6, 16, 2)                              # one sample plane only


@pytest.mark.gpu
def test_device_slab_meshes_concatenate_to_the_full_mesh():
    import torch
    from dynamicfusion_body_b200 import engine
    for name, vol, step, level, world in _slab_cases():
        full = engine.marching_cubes(torch.from_numpy(vol).cuda(), step, level)
        got = _slabwise(lambda sub, s, lv, **kw: engine.marching_cubes(sub.cuda(), s, lv, **kw), vol, step, level, world)
        _assert_identical(got, full, name)
        assert_same_mesh(full, omc.marching_cubes(vol, step, level), name)


@pytest.mark.gpu
def test_fusion_classes_extract_on_the_device(tmp_path):
    """Fusion.marching_cubes / InitializeCanonicalSpace / write_canonical_mesh with the default ("device") extractor; a lone x-slab
    comes out in whole-grid coordinates."""
    from dynamicfusion_body_b200 import fusion
    vol = np.clip(sphere(32, (15.2, 16.1, 15.7), 10.3), -3, 3).astype(np.float32)
    want = omc.marching_cubes(vol, 1, None)
    fus = fusion.Fusion(3.0, subsample_rate=2.0, knn=3, marching_cubes_step_size=1, verbose=False, use_cnn=False, write_warpfield=False)
    fus.InitializeCanonicalSpace(tsdf=vol)                                  # core/fusion.py:73-96: marching cubes -> radius -> graph
    assert np.array_equal(fus._vertices, want[0]) and np.array_equal(fus._faces, want[1])
    assert np.allclose(fus._normals, want[2], atol=1e-6) and len(fus._nodes) > 5 and fus._radius > 1.0
    fus.write_canonical_mesh(str(tmp_path), "canonical.obj")
    lines = open(os.path.join(str(tmp_path), "canonical.obj")).read().split("\n")
    assert sum(l.startswith("v ") for l in lines) == len(want[0]) and sum(l.startswith("f ") for l in lines) == len(want[1])
    slab = fusion.Fusion(3.0, subsample_rate=2.0, knn=3, marching_cubes_step_size=1, verbose=False, use_cnn=False, write_warpfield=False)
    slab._set_volume(vol[8:20].copy(), None, shape=(32, 32, 32), slab=(8, 20))
    slab.marching_cubes()
    assert_same_mesh((slab._vertices, slab._faces, slab._normals, want[3][:0]), omc.marching_cubes(vol[8:20], 1, None, x_origin=8)[:3] + (want[3][:0],), "slab")
