"""Synthetic, seeded inputs for tests and benchmarks (host-side numpy; input synthesis, not the hot path).

The reference ships no data (`data/` is git-ignored, core/__init__.py:8); its only asset is the canonical
body mesh `meshes/original.obj` (16 741 vertices in [0,64]^3 voxel coordinates).  A re-encoded copy of
that mesh (float32 vertices/normals, int32 faces) is committed as tests/golden/body_mesh.npz by
tests/golden/make_golden.py; everything else -- deformation nodes, a smooth dual-quaternion warp field,
cameras, z-buffered depth maps of the warped mesh, live TSDF volumes -- is generated here from a seed.
"""
import dataclasses
import os

import numpy as np
from scipy.spatial import cKDTree

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MESH_FIXTURE = os.path.join(_REPO, "tests", "golden", "body_mesh.npz")


def load_body_mesh(path=None):
    """(vertices f32 (V,3) in [0,64]^3, normals f32 (V,3), faces i32 (F,3) 0-based)."""
    path = path or MESH_FIXTURE
    if os.path.isfile(path):
        z = np.load(path)
        return z["vertices"].astype(np.float32), z["normals"].astype(np.float32), z["faces"].astype(np.int32)
    # procedural stand-in (an ellipsoid "torso") so that nothing hard-fails without the fixture
    n = 16000
    i = np.arange(n) + 0.5
    phi = np.arccos(1 - 2 * i / n)
    th = np.pi * (1 + 5 ** 0.5) * i
    d = np.stack([np.cos(th) * np.sin(phi), np.sin(th) * np.sin(phi), np.cos(phi)], 1)
    v = (d * np.array([12.0, 20.0, 28.0]) + 32.0).astype(np.float32)
    nr = d / np.array([12.0, 20.0, 28.0])
    nr = (nr / np.linalg.norm(nr, axis=1, keepdims=True)).astype(np.float32)
    return v, nr, np.zeros((0, 3), np.int32)


def uniform_sample(points, radius):
    """Radius-based greedy subsampling with the semantics of the reference's `uniform_sample`
    (core/util.py:27-47): walk the points in order, keep the first one still alive, drop every point
    strictly closer than `radius` to it.  KD-tree accelerated; returns (samples, indices)."""
    pts = np.asarray(points)
    p64 = pts.astype(np.float64)
    tree = cKDTree(p64)
    alive = np.ones(len(pts), dtype=bool)
    keep = []
    for i in range(len(pts)):
        if not alive[i]:
            continue
        keep.append(i)
        nb = np.asarray(tree.query_ball_point(p64[i], radius), dtype=np.int64)
        d = np.linalg.norm(p64[nb] - p64[i], axis=1)
        alive[nb[d < radius]] = False
    keep = np.asarray(keep, dtype=np.int64)
    return pts[keep].copy(), keep


def se3_to_dq(Rm, t):
    """Unit dual quaternion [q, 0.5 * (0,t) * q] of x -> Rm x + t (convention of core/util.py:79-84)."""
    tr = np.trace(Rm)
    if tr > 0:
        s = np.sqrt(tr + 1.0) * 2
        q = np.array([0.25 * s, (Rm[2, 1] - Rm[1, 2]) / s, (Rm[0, 2] - Rm[2, 0]) / s, (Rm[1, 0] - Rm[0, 1]) / s])
    else:
        i = int(np.argmax(np.diag(Rm)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(1.0 + Rm[i, i] - Rm[j, j] - Rm[k, k]) * 2
        q = np.zeros(4)
        q[0] = (Rm[k, j] - Rm[j, k]) / s
        q[1 + i] = 0.25 * s
        q[1 + j] = (Rm[j, i] + Rm[i, j]) / s
        q[1 + k] = (Rm[k, i] + Rm[i, k]) / s
    q = q / np.linalg.norm(q)
    return np.concatenate([q, 0.5 * _qmul(np.array([0.0, t[0], t[1], t[2]]), q)])


def _qmul(a, b):
    w1, x1, y1, z1 = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    w0, x0, y0, z0 = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack([w1 * w0 - x1 * x0 - y1 * y0 - z1 * z0,
                     w1 * x0 + x1 * w0 + y1 * z0 - z1 * y0,
                     w1 * y0 - x1 * z0 + y1 * w0 + z1 * x0,
                     w1 * z0 + x1 * y0 - y1 * x0 + z1 * w0], -1)


def axis_angle_dq(axis, angle, trans):
    """Batched unit dual quaternions: rotation (axis, angle) about the ORIGIN followed by `trans`."""
    axis = axis / np.linalg.norm(axis, axis=-1, keepdims=True)
    q = np.concatenate([np.cos(angle / 2)[..., None], np.sin(angle / 2)[..., None] * axis], -1)
    tq = np.concatenate([np.zeros(trans.shape[:-1] + (1,)), trans], -1)
    return np.concatenate([q, 0.5 * _qmul(tq, q)], -1)


def dq_apply(dq, p):
    """Closed form of the reference's dqb_warp for (possibly non-unit) dq: |r|^2 R(r) p + t."""
    w, v, dw, dv = dq[..., 0:1], dq[..., 1:4], dq[..., 4:5], dq[..., 5:8]
    rot = (w * w - (v * v).sum(-1, keepdims=True)) * p + 2 * (v * p).sum(-1, keepdims=True) * v + 2 * w * np.cross(v, p)
    return rot + 2 * (w * dv - dw * v + np.cross(v, dv))


def smooth_warp_field(node_pos, rng, max_disp=0.3, extent=64.0):
    """Smooth random SE(3) per node as unit dual quaternions (N,8) float32: a small rotation about the
    node itself plus a small translation, both low-frequency functions of position.

    The magnitudes are bounded on purpose.  The reference normalises the blended dual quaternion by its
    8-vector norm (Q2, core/fusion.py:551), so a blend whose dual part is not << 1 -- i.e. a node
    translation t = c - R c + d of more than a fraction of a voxel -- contracts space by
    1/(1 + |t|^2/4).  `max_disp` (voxels) bounds |t| so that the synthetic live frame stays a plausible
    small inter-frame motion under the reference's own arithmetic."""
    p = node_pos.astype(np.float64)
    n = len(p)
    f = rng.uniform(0.5, 1.5, size=(7, 3)) * (2 * np.pi / extent)
    ph = rng.uniform(0, 2 * np.pi, size=7)
    fields = np.sin(p @ f.T + ph)                       # (N,7) smooth in space
    axis = fields[:, 0:3] + 1e-3
    angle = 0.5 * max_disp / np.maximum(np.linalg.norm(p, axis=1), 1.0) * 0.5 * (1 + fields[:, 3])
    d = 0.5 * max_disp * fields[:, 4:7]
    return axis_angle_dq(axis, angle, d).astype(np.float32)


def blend_warp(points, node_pos, node_dq, node_w, knn_idx, lw=None):
    """Plain float64 DQB warp (8-norm normalisation, no float32 rounding games) used to synthesise the
    live-frame geometry.  This is input synthesis only -- parity is judged against oracle/."""
    p = points.astype(np.float64)
    npk = node_pos.astype(np.float64)[knn_idx]
    d2 = ((p[:, None, :] - npk) ** 2).sum(-1)
    w = np.exp(-d2 / (4.0 * np.asarray(node_w, dtype=np.float64)[knn_idx] ** 2))
    b = (w[..., None] * node_dq.astype(np.float64)[knn_idx]).sum(1)
    b /= np.linalg.norm(b, axis=1, keepdims=True)
    out = dq_apply(b, p)
    if lw is not None:
        out = dq_apply(np.asarray(lw, dtype=np.float64)[None, :], out)
    return out


def look_at_extrinsic(eye, target, up=(0.0, 1.0, 0.0)):
    """3x4 world->camera with +z looking from `eye` to `target`."""
    eye = np.asarray(eye, dtype=np.float64)
    zc = np.asarray(target, dtype=np.float64) - eye
    zc /= np.linalg.norm(zc)
    xc = np.cross(np.asarray(up, dtype=np.float64), zc)
    if np.linalg.norm(xc) < 1e-6:
        xc = np.cross(np.array([1.0, 0, 0]), zc)
    xc /= np.linalg.norm(xc)
    yc = np.cross(zc, xc)
    Rm = np.stack([xc, yc, zc], 0)
    return np.concatenate([Rm, (-Rm @ eye)[:, None]], 1)


def render_depth(verts_cam, faces, K, rows, cols):
    """z-buffer of a triangle mesh given in CAMERA coordinates; returns the reference's depth-map
    convention (core/fusion_dm.py:196): depth stored NEGATIVE, 0 = no data.  Triangles are densely
    point-sampled (spacing < 1/2 pixel) and splatted with a min-reduction."""
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    if len(faces) == 0:
        pts = verts_cam
    else:
        a, b, c = verts_cam[faces[:, 0]], verts_cam[faces[:, 1]], verts_cam[faces[:, 2]]
        zmin = max(1e-6, float(np.percentile(np.concatenate([a[:, 2], b[:, 2], c[:, 2]]), 1)))
        edge = max(np.linalg.norm(a - b, axis=1).max(), np.linalg.norm(b - c, axis=1).max(), np.linalg.norm(c - a, axis=1).max())
        s = int(min(24, max(1, np.ceil(edge * max(fx, fy) / zmin / 0.5))))
        i, j = np.meshgrid(np.arange(s + 1), np.arange(s + 1), indexing="ij")
        m = (i + j) <= s
        u, v = i[m] / s, j[m] / s
        w = 1.0 - u - v
        pts = (w[None, :, None] * a[:, None, :] + u[None, :, None] * b[:, None, :] + v[None, :, None] * c[:, None, :]).reshape(-1, 3)
    z = pts[:, 2]
    ok = z > 1e-6
    pts, z = pts[ok], z[ok]
    px = fx * pts[:, 0] / z + cx
    py = fy * pts[:, 1] / z + cy
    buf = np.full(rows * cols, np.inf)
    for ox in (0, 1):
        for oy in (0, 1):
            ui = np.floor(px).astype(np.int64) + ox
            vi = np.floor(py).astype(np.int64) + oy
            inside = (ui >= 0) & (ui < cols) & (vi >= 0) & (vi < rows)
            np.minimum.at(buf, vi[inside] * cols + ui[inside], z[inside])
    dm = np.where(np.isfinite(buf), -buf, 0.0).reshape(rows, cols)
    return dm.astype(np.float32)


def mesh_sdf_volume(shape, verts, normals, trunc=None, chunk=1 << 20):
    """Point-to-plane signed distance to the nearest mesh vertex on the integer grid of `shape`
    (positive outside).  Untruncated unless `trunc` is given (then clipped to [-trunc, trunc])."""
    tree = cKDTree(verts.astype(np.float64))
    n = int(np.prod(shape))
    out = np.empty(n, dtype=np.float32)
    idx = np.arange(n)
    for s in range(0, n, chunk):
        ii = idx[s:s + chunk]
        g = np.stack(np.unravel_index(ii, shape), 1).astype(np.float64)
        _, nn = tree.query(g, workers=-1)
        out[s:s + chunk] = ((g - verts[nn]) * normals[nn]).sum(1)
    out = out.reshape(shape)
    if trunc is not None:
        out = np.clip(out, -trunc, trunc)
    return out


@dataclasses.dataclass
class Scene:
    res: int
    k: int
    vertices: np.ndarray        # (V,3) f32 canonical surface vertices, grid coordinates
    normals: np.ndarray         # (V,3) f32
    faces: np.ndarray           # (F,3) i32
    node_pos: np.ndarray        # (N,3) f32
    node_idx: np.ndarray        # (N,)  vertex index of each node
    node_dq: np.ndarray         # (N,8) f32
    node_w: float               # dg_w = 2*radius for every node (core/fusion.py:116)
    radius: float
    lw: np.ndarray              # (8,) global rigid dq
    K: np.ndarray               # (3,3)
    Kinv: np.ndarray
    rows: int
    cols: int
    extrinsics: object          # None or (n_views,3,4)
    depths: np.ndarray          # (n_views,rows,cols) f32, negative depth
    tdist: float
    vert_knn: np.ndarray        # (V,k) node ids per vertex (Fusion._neighbor_look_up)
    warped_vertices: np.ndarray  # (V,3) f64 live-frame vertices (lw applied)
    warped_normals: np.ndarray

    @property
    def n_nodes(self):
        return len(self.node_pos)

    def nodes_as_reference_tuples(self):
        """`Fusion._nodes` layout, core/fusion.py:113-116."""
        return [(int(self.node_idx[i]), self.node_pos[i], self.node_dq[i], float(self.node_w)) for i in range(self.n_nodes)]


def make_scene(res=64, k=4, radius=None, n_nodes=None, seed=0, n_views=1, rows=480, cols=640, tdist=None,
               max_disp=0.3, lw_dtype=np.float64, unit_init=False, mesh_path=None,
               focal=None, cam_dist=1.7, background=False, view_axis="z"):
    """Seeded benchmark / test scene.  `radius` in units of the 64^3 mesh; either `radius` or
    `n_nodes` (bisection on the radius) may be given.  One view: the global rigid dq `lw` is the camera
    extrinsic; several views: `lw` is a small rigid motion and cameras sit on a ring (extrinsics).  `view_axis` (one view): the
    grid axis the camera looks along -- "z" is the long axis of the update pass's 4x4x32 bricks, "x" / "y" look across them."""
    rng = np.random.default_rng(seed)
    v64, nrm, faces = load_body_mesh(mesh_path)
    scale = (res - 1) / 64.0
    verts = (v64 * scale).astype(np.float32)
    if radius is None:
        lo, hi = 0.5, 16.0
        target = n_nodes or 1000
        for _ in range(18):
            mid = 0.5 * (lo + hi)
            cnt = len(uniform_sample(v64, mid)[1])
            if cnt > target:
                lo = mid
            else:
                hi = mid
            if abs(cnt - target) <= max(2, target // 100):
                break
        radius = mid
    node64, node_idx = uniform_sample(v64, radius)
    node_pos = verts[node_idx].copy()
    r_grid = float(radius * scale)
    node_w = 2.0 * r_grid
    if unit_init:
        node_dq = np.tile(np.array([1, 0, 0, 0, 0, 0.01, 0.01, 0], dtype=np.float32), (len(node_pos), 1))  # Q5
    else:
        node_dq = smooth_warp_field(node_pos, rng, max_disp, extent=float(res))
    tree = cKDTree(node_pos.astype(np.float64))
    _, vert_knn = tree.query(verts.astype(np.float64), k=k)
    vert_knn = np.atleast_2d(vert_knn).reshape(len(verts), k).astype(np.int64)

    centre = np.array([(res - 1) / 2.0] * 3)
    focal = focal if focal is not None else 525.0 * cols / 640.0
    K = np.array([[focal, 0, (cols - 1) / 2.0], [0, focal, (rows - 1) / 2.0], [0, 0, 1.0]])
    Kinv = np.linalg.inv(K)
    dist = cam_dist * res
    if n_views == 1:
        off = {"z": np.array([0.15 * res, -0.1 * res, -dist]), "x": np.array([-dist, -0.1 * res, 0.15 * res]),
               "y": np.array([0.15 * res, -dist, -0.1 * res])}[view_axis]
        E = look_at_extrinsic(centre + off, centre, up=(0.0, 0.0, 1.0) if view_axis == "y" else (0.0, 1.0, 0.0))
        lw = se3_to_dq(E[:, :3], E[:, 3])
        extr = None
    else:
        ang = np.deg2rad(1.5)
        Rl = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1.0]])
        lw = se3_to_dq(Rl, centre - Rl @ centre + np.array([0.3, -0.2, 0.1]) * scale)
        extr = np.stack([look_at_extrinsic(centre + dist * np.array([np.sin(a), 0.1, -np.cos(a)]), centre)
                         for a in np.arange(n_views) * 2 * np.pi / n_views])
    lw = lw.astype(lw_dtype)
    wv = blend_warp(verts, node_pos, node_dq, np.full(len(node_pos), node_w), vert_knn, lw=lw.astype(np.float64))
    # warped normals: rotate with the blended real part (scaled like the reference; renormalised for synthesis)
    eps = 1e-3
    wv2 = blend_warp(verts + eps * nrm, node_pos, node_dq, np.full(len(node_pos), node_w), vert_knn, lw=lw.astype(np.float64))
    wn = wv2 - wv
    wn /= np.maximum(np.linalg.norm(wn, axis=1, keepdims=True), 1e-12)
    depths = []
    for vi in range(n_views):
        vc = wv if extr is None else wv @ extr[vi][:, :3].T + extr[vi][:, 3]
        dm = render_depth(vc, faces, K, rows, cols)
        if background:
            # a wall behind the capture volume: every pixel carries a measurement, as in a real sensor
            dm[dm == 0] = -np.float32(dist + 0.9 * res)
        depths.append(dm)
    if tdist is None:
        tdist = 5.0 * res / 256.0  # test.py:159: 0.2 world units in a volume 8*std = 10.4 units wide (fusion_dm.py:107,136)
    return Scene(res=res, k=k, vertices=verts, normals=nrm, faces=faces, node_pos=node_pos, node_idx=node_idx,
                 node_dq=node_dq, node_w=node_w, radius=r_grid, lw=lw, K=K, Kinv=Kinv, rows=rows, cols=cols,
                 extrinsics=extr, depths=np.stack(depths), tdist=float(tdist), vert_knn=vert_knn,
                 warped_vertices=wv, warped_normals=wn)


@dataclasses.dataclass
class GNProblemData:
    vertices: np.ndarray      # (V,3) f32 canonical surface samples
    normals: np.ndarray       # (V,3) f32
    corr: np.ndarray          # (V,3) f64 live-frame correspondences
    vert_knn: np.ndarray      # (V,k) int64
    node_vertex_idx: np.ndarray
    x0: np.ndarray            # (8N,) f64 initial node transforms (perturbed truth)
    x_true: np.ndarray


def make_gn_problem(sc, n_points, seed=0, noise=0.02, perturb=2e-3):
    """Solver workload (BASELINE config 3 shape): `n_points` area-weighted samples of the canonical mesh as
    (vertex, normal, correspondence) triples -- the stand-in for the valid pixels of one depth frame -- with
    correspondences produced by the scene's true warp field (+ noise) and a perturbed initial field."""
    rng = np.random.default_rng(seed)
    v = sc.vertices.astype(np.float64)
    f = sc.faces
    if len(f) == 0:
        idx = rng.integers(0, len(v), n_points)
        pts, nrm = v[idx], sc.normals[idx].astype(np.float64)
    else:
        a, b, c = v[f[:, 0]], v[f[:, 1]], v[f[:, 2]]
        area = 0.5 * np.linalg.norm(np.cross(b - a, c - a), axis=1)
        tri = rng.choice(len(f), size=n_points, p=area / area.sum())
        r1, r2 = rng.random(n_points), rng.random(n_points)
        s1 = np.sqrt(r1)
        w0, w1, w2 = 1 - s1, s1 * (1 - r2), s1 * r2
        pts = w0[:, None] * a[tri] + w1[:, None] * b[tri] + w2[:, None] * c[tri]
        n0 = sc.normals.astype(np.float64)
        nrm = w0[:, None] * n0[f[tri, 0]] + w1[:, None] * n0[f[tri, 1]] + w2[:, None] * n0[f[tri, 2]]
        nrm /= np.maximum(np.linalg.norm(nrm, axis=1, keepdims=True), 1e-12)
    pts32 = pts.astype(np.float32)
    tree = cKDTree(sc.node_pos.astype(np.float64))
    _, knn = tree.query(pts32.astype(np.float64), k=sc.k)
    knn = knn.reshape(len(pts32), sc.k).astype(np.int64)
    nw = np.full(sc.n_nodes, np.float32(sc.node_w))
    corr = blend_warp(pts32, sc.node_pos, sc.node_dq, nw, knn, lw=sc.lw.astype(np.float64)) + rng.normal(size=pts.shape) * noise
    _, nvi = cKDTree(pts32.astype(np.float64)).query(sc.node_pos.astype(np.float64))
    x_true = sc.node_dq.reshape(-1).astype(np.float64)
    x0 = x_true + rng.normal(size=x_true.shape) * perturb
    return GNProblemData(vertices=pts32, normals=nrm.astype(np.float32), corr=corr, vert_knn=knn, node_vertex_idx=nvi.astype(np.int64),
                         x0=x0, x_true=x_true)
