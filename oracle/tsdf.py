"""CPU oracle, part 2: the TSDF update paths -- TEST INFRASTRUCTURE ONLY (see oracle/dq.py header).

a1  update_volume      = Fusion.updateTSDF        core/fusion.py:153-198
a2  fuse_depth_rigid   = FusionDM.fuseDepths      core/fusion_dm.py:180-217
a3  update_projective  = Fusion.warp (core/fusion.py:502-520) composed with the per-voxel body of
                         FusionDM.fuseDepths (core/fusion_dm.py:193-210) -- the north-star path
    update_rigid_volume= FusionDM.updateTSDF      core/fusion_dm.py:300-316

All functions are vectorised over a list of voxel multi-indices so they can be run on whole grids,
slabs, or random voxel subsets, and return masks alongside the updated values.
"""
import numpy as np

from . import dq as _dq


def voxel_grid(shape, x0=0, x1=None):
    """Multi-indices in np.nditer C order (x slowest, z fastest; core/fusion.py:171) as float32
    (`np.array(it.multi_index, dtype=np.float32)`, core/fusion.py:174)."""
    x1 = shape[0] if x1 is None else x1
    g = np.indices((x1 - x0, shape[1], shape[2]), dtype=np.int32).reshape(3, -1).T.copy()
    g[:, 0] += x0
    return g.astype(np.float32)


def interpolate_tsdf(pos, tsdf):
    """core/util.py:102-137, vectorised.  Returns (value float64, valid bool).
    Q1: after the x-lerp the reference combines (c00,c10) with yd, (c01,c11) with yd and then the two
    with zd, where c001/c101 are the **y1** corners and c010/c110 the **z1** corners -- i.e. the y/z
    fractional weights are swapped relative to textbook trilinear.  Reproduced literally."""
    if tsdf.ndim != 3:
        raise ValueError('Only 3D numpy array is accepted')
    pos = np.asarray(pos)
    rx, ry, rz = tsdf.shape
    valid = ~((pos.min(axis=-1) < 0) | (pos[..., 0] > rx - 1) | (pos[..., 1] > ry - 1) | (pos[..., 2] > rz - 1))
    valid &= np.isfinite(pos).all(axis=-1)
    p = np.where(valid[..., None], pos, 0.0)
    f = np.floor(p)
    c = np.ceil(p)
    x0, y0, z0 = (f[..., i].astype(np.int64) for i in range(3))
    x1, y1, z1 = (c[..., i].astype(np.int64) for i in range(3))
    xd = p[..., 0] - f[..., 0]
    yd = p[..., 1] - f[..., 1]
    zd = p[..., 2] - f[..., 2]
    c000 = tsdf[x0, y0, z0]
    c100 = tsdf[x1, y0, z0]
    c001 = tsdf[x0, y1, z0]
    c101 = tsdf[x1, y1, z0]
    c010 = tsdf[x0, y0, z1]
    c110 = tsdf[x1, y0, z1]
    c011 = tsdf[x0, y1, z1]
    c111 = tsdf[x1, y1, z1]
    c00 = c000 * (1 - xd) + c100 * xd
    c01 = c001 * (1 - xd) + c101 * xd
    c10 = c010 * (1 - xd) + c110 * xd
    c11 = c011 * (1 - xd) + c111 * xd
    c0 = c00 * (1 - yd) + c10 * yd
    c1 = c01 * (1 - yd) + c11 * yd
    return c0 * (1 - zd) + c1 * zd, valid


def _gather_nodes(knn_idx, node_pos, node_dq, node_w):
    return node_pos[knn_idx], node_dq[knn_idx], np.asarray(node_w, dtype=np.float64)[knn_idx]


def mean_node_distance(pos, node_pos_k):
    """Q4 weight, core/fusion.py:180-183: `wi += la.norm(nodes[idx][1] - pos) / len(locations)`
    starting from python int 0; each norm is float32 when both operands are float32, the division by
    the python int k keeps float32, and the running sum is float32 as well."""
    k = node_pos_k.shape[-2]
    wi = None
    for i in range(k):
        term = _dq.norm3_like_la(node_pos_k[..., i, :] - pos) / k
        wi = term if wi is None else wi + term
    return wi


def update_volume(tsdf, tsdfw, curr_tsdf, vox, knn_idx, node_pos, node_dq, node_w, lw, tdist, wmax=100.0):
    """a1: Fusion.updateTSDF body (core/fusion.py:171-190) for voxels `vox` (M,3) float32 with
    their k nearest nodes `knn_idx` (M,k).  tsdf/tsdfw are flat views aligned with `vox`.
    Returns (new_tsdf, new_w, mask) as float64/float64/bool of length M; inputs are not modified."""
    npk, ndk, nwk = _gather_nodes(knn_idx, node_pos, node_dq, node_w)
    pw = _dq.warp(vox, npk, ndk, nwk, lw=lw)
    tl, valid = interpolate_tsdf(pw, curr_tsdf)
    with np.errstate(invalid='ignore'):
        mask = valid & (tl > -1 * tdist)
    wi_n = mean_node_distance(vox, npk)          # float32 when vox and node_pos are float32 (Q4)
    wi = wi_n.astype(np.float64)
    v_old = np.asarray(tsdf)
    w_old = np.asarray(tsdfw)
    wi_t = np.where(w_old == 0, wi, w_old)
    with np.errstate(invalid='ignore', divide='ignore', over='ignore'):
        # `min(self._tdist, tsdf_l) * wi`: python min() returns the python-float tdist (a weak scalar,
        # so the product is rounded in wi's dtype) unless tsdf_l < tdist (np.float64, strong).
        clamped = (wi_n.dtype.type(tdist) * wi_n).astype(np.float64)
        term = np.where(tl < tdist, tl * wi, clamped)
        v_new = (v_old * wi_t + term) / (wi + wi_t)
        w_new = np.minimum(wi + wi_t, wmax)
    return (np.where(mask, v_new, v_old).astype(np.float64),
            np.where(mask, w_new, w_old).astype(np.float64), mask)


def _matvec_rows(M, cols):
    """np.matmul(M, v) for ONE small vector v, vectorised over many v: numpy evaluates each output element as the plain
    left-to-right sum  M[r,0]*v0 + M[r,1]*v1 + ...  (no FMA, no blocking).  A batched `V @ M.T` goes through dgemm, whose
    different rounding flips knife-edge decisions (e.g. u == k + 0.5 exactly when the optical axis passes through voxel
    centres) -- caught by tests/test_gpu_classes.py against the reference."""
    M = np.asarray(M, dtype=np.float64)
    out = []
    for r in range(M.shape[0]):
        acc = M[r, 0] * cols[0]
        for c in range(1, M.shape[1]):
            acc = acc + M[r, c] * cols[c]
        out.append(acc)
    return np.stack(out, axis=-1)


def _project_and_fuse(lpos, dm, K, Kinv, v_old, w_old, tdist, scale, wmax):
    """Per-voxel body of FusionDM.fuseDepths from `project_to_pixel` on (core/fusion_dm.py:194-210,
    core/util.py:312-320).  `(dmx, dmy) = dm.shape` = (rows, cols): u is tested against cols-1 and v
    against rows-1; the pixel is `dm[int(round(v))][int(round(u))]` (round-half-even); depth is stored
    negative; `z > 0` required."""
    rows, cols = dm.shape
    p = _matvec_rows(K, [lpos[..., 0], lpos[..., 1], lpos[..., 2]])
    nz = p[..., 2] != 0
    with np.errstate(invalid='ignore', divide='ignore'):
        u = p[..., 0] / p[..., 2]
        v = p[..., 1] / p[..., 2]
        frustum = nz & (u >= 0) & (u < cols - 1) & (v >= 0) & (v < rows - 1)
    ui = np.where(frustum, np.rint(np.where(frustum, u, 0)), 0).astype(np.int64)
    vi = np.where(frustum, np.rint(np.where(frustum, v, 0)), 0).astype(np.int64)
    z = -1 * dm[vi, ui]
    has_depth = frustum & (z > 0)
    uc = z[..., None] * np.stack([u, v, np.ones_like(u)], axis=-1)
    with np.errstate(invalid='ignore'):
        cpos = _matvec_rows(Kinv, [uc[..., 0], uc[..., 1], uc[..., 2]])
        tl = cpos[..., 2] - lpos[..., 2]
        mask = has_depth & (tl > -1 * tdist)
    wi = 1
    with np.errstate(invalid='ignore', divide='ignore'):
        v_new = (scale * v_old * w_old + np.minimum(tdist, tl) * wi) / (scale * (wi + w_old))
        w_new = np.minimum(wi + w_old, wmax)
    return (np.where(mask, v_new, v_old).astype(np.float64),
            np.where(mask, w_new, w_old).astype(np.float64), mask, frustum)


def fuse_depth_rigid(tsdf, tsdfw, vox, dm, lw34, K, Kinv, tdist, res, scale=1.0, center=None, wmax=100.0):
    """a2: FusionDM.fuseDepths (core/fusion_dm.py:180-217) for voxels `vox` (M,3) float32.
    `pos = scale * (pos - tsdf_res/2) + center` (:183,:191) in float64, `lpos = lw @ [pos,1]` (:193).
    Returns (new_tsdf, new_w, mask, frustum)."""
    center = np.zeros(3) if center is None else center
    sdf_center = np.zeros(3) + res / 2
    pos = scale * (vox - sdf_center) + center
    lw34 = np.asarray(lw34)
    lpos = _matvec_rows(lw34, [pos[..., 0], pos[..., 1], pos[..., 2], np.ones(pos.shape[:-1])])
    return _project_and_fuse(lpos, dm, K, Kinv, np.asarray(tsdf), np.asarray(tsdfw), tdist, scale, wmax)


def update_projective(tsdf, tsdfw, vox, knn_idx, node_pos, node_dq, node_w, lw, dms, K, Kinv, tdist,
                      extrinsics=None, wmax=100.0):
    """a3 (north-star composition): `p' = Fusion.warp(pos, dqs, locations,
    m_lw=lw)` (core/fusion.py:502-520), `lpos = E_view @ [p',1]` (or p' when no extrinsic), then the
    fuseDepths body with scale=1, center=0; views are applied sequentially in index order
    (core/fusion_dm.py:166-170).  Returns (new_tsdf, new_w, masks (V,M), frusta (V,M))."""
    npk, ndk, nwk = _gather_nodes(knn_idx, node_pos, node_dq, node_w)
    pw = _dq.warp(vox, npk, ndk, nwk, lw=lw)
    v = np.asarray(tsdf, dtype=np.float64)
    w = np.asarray(tsdfw, dtype=np.float64)
    masks, frusta = [], []
    for vi, dm in enumerate(dms):
        if extrinsics is not None:
            E = np.asarray(extrinsics[vi], dtype=np.float64)
            lpos = _matvec_rows(E, [pw[..., 0], pw[..., 1], pw[..., 2], np.ones(pw.shape[:-1])])
        else:
            lpos = pw
        v, w, m, fr = _project_and_fuse(lpos, dm, K, Kinv, v, w, tdist, 1.0, wmax)
        masks.append(m)
        frusta.append(fr)
    return v, w, np.array(masks), np.array(frusta)


def update_rigid_volume(tsdf, tsdfw, curr_tsdf, vox, lw, tdist, wmax=100.0):
    """FusionDM.updateTSDF (core/fusion_dm.py:300-316): global rigid dq only, unit weight."""
    pw = _dq.dqb_warp(lw, vox)
    tl, valid = interpolate_tsdf(pw, curr_tsdf)
    with np.errstate(invalid='ignore'):
        mask = valid & (tl > -1 * tdist)
    v_old = np.asarray(tsdf)
    w_old = np.asarray(tsdfw)
    with np.errstate(invalid='ignore', divide='ignore'):
        v_new = (v_old * w_old + np.minimum(tdist, tl) * 1) / (1 + w_old)
        w_new = np.minimum(1 + w_old, wmax)
    return (np.where(mask, v_new, v_old).astype(np.float64),
            np.where(mask, w_new, w_old).astype(np.float64), mask)
