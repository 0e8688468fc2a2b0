"""On-disk formats of the reference so that its datasets run through the new path unchanged.
Host-side numpy; not part of the hot path."""
import os
import pickle

import numpy as np


def load_sdf(file_path, read_closest_points=False, verbose=False):
    """`.dist` signed-distance file (format: core/sdf.py:10-21; reader semantics: core/sdf.py:24-69).
    Header: three int32 resolutions -- the first two are stored NEGATED -- then b_min, b_max (3 float64 each), then
    (rx+1)(ry+1)(rz+1) float32 distances with x fastest; optionally 3 float32 closest-point coordinates per grid vertex.
    Returns (b_min, b_max, volume[x][y][z] float32, closest_points or None) like the reference."""
    with open(file_path, 'rb') as fp:
        res = np.fromfile(fp, dtype=np.int32, count=3)
        if res.size != 3:
            raise ValueError('truncated .dist header')
        res_x, res_y, res_z = -int(res[0]), -int(res[1]), int(res[2])
        if min(res_x, res_y, res_z) < 0:
            raise ValueError('bad .dist resolution header %s' % (res,))
        if verbose:
            print("resolution: %d %d %d" % (res_x, res_y, res_z))
        b_min = np.fromfile(fp, dtype=np.float64, count=3)
        b_max = np.fromfile(fp, dtype=np.float64, count=3)
        grid_num = (1 + res_x) * (1 + res_y) * (1 + res_z)
        volume = np.fromfile(fp, dtype=np.float32, count=grid_num)
        if volume.size != grid_num:
            raise ValueError('truncated .dist distance block')
        volume = np.swapaxes(volume.reshape((1 + res_z, 1 + res_y, 1 + res_x)), 0, 2)
        closest_points = None
        if read_closest_points:
            cp = np.fromfile(fp, dtype=np.float32, count=grid_num * 3)
            if cp.size != grid_num * 3:
                raise ValueError('truncated .dist closest-point block')
            closest_points = np.swapaxes(cp.reshape((1 + res_z, 1 + res_y, 1 + res_x, 3)), 0, 2)
    return b_min, b_max, volume, closest_points


def save_sdf(file_path, volume, b_min=(0, 0, 0), b_max=None, closest_points=None):
    """Writer for the same format (the reference has none); used to build test fixtures and synthetic sequences."""
    volume = np.asarray(volume, dtype=np.float32)
    rx, ry, rz = (s - 1 for s in volume.shape)
    b_max = b_max if b_max is not None else (rx, ry, rz)
    with open(file_path, 'wb') as fp:
        np.array([-rx, -ry, rz], dtype=np.int32).tofile(fp)
        np.asarray(b_min, dtype=np.float64).tofile(fp)
        np.asarray(b_max, dtype=np.float64).tofile(fp)
        np.ascontiguousarray(np.swapaxes(volume, 0, 2)).tofile(fp)
        if closest_points is not None:
            np.ascontiguousarray(np.swapaxes(np.asarray(closest_points, dtype=np.float32), 0, 2)).tofile(fp)


def read_proj_matrix(fpath):
    """Whitespace-separated projection matrix text file (core/util.py:330-335); test.py:155 turns it into the 3x4
    extrinsic with `Kinv @ P`."""
    with open(fpath, 'r') as f:
        return np.array([line.split() for line in f if line.strip()], dtype='float')


def write_warp_field(nodes, path, filename, itercounter):
    """core/fusion.py:571-573: pickle of the `_nodes` list to <path>/<filename>__<iter>.p."""
    with open(os.path.join(path, filename + '__' + str(itercounter) + '.p'), 'wb') as f:
        pickle.dump(list(nodes), f)


def read_warp_field(fpath):
    with open(fpath, 'rb') as f:
        return pickle.load(f)


def write_obj(fpath, verts, normals=None, faces=None, face_normals=False):
    """OBJ text in the reference's two flavours: Fusion.write_canonical_mesh (core/fusion.py:577-586: `v`, `vn`, `f a b c`)
    and FusionDM.write_canonical_mesh (core/fusion_dm.py:339-354: faces as `f a//a b//b c//c`); 1-based indices, %f formatting."""
    def rows(fmt, a, cols):
        # one C-level format call per 64 k rows instead of a Python-level write per row (605 k vertices + 1.2 M faces at 512^3);
        # '%f' / '%d' see the same Python floats / ints as the reference's per-row '%' does, so the text is byte-identical
        a = np.asarray(a).reshape(-1, cols)
        for i in range(0, len(a), 65536):
            blk = a[i:i + 65536]
            yield (fmt * len(blk)) % tuple(blk.ravel().tolist())

    with open(fpath, 'w') as f:
        f.writelines(rows('v %f %f %f\n', verts, 3))
        if normals is not None:
            f.writelines(rows('vn %f %f %f\n', normals, 3))
        if faces is not None and len(faces):
            t = np.asarray(faces).reshape(-1, 3).astype(np.int64) + 1
            if face_normals:
                f.writelines(rows('f %d//%d %d//%d %d//%d\n', np.repeat(t, 2, axis=1), 6))
            else:
                f.writelines(rows('f %d %d %d\n', t, 3))
