"""TEST INFRASTRUCTURE ONLY -- CPU statement of the surface extractor that the CUDA path
(dynamicfusion_body_b200/csrc/mc.cu) is checked against.  Never imported by the product.

PARITY UNPINNED against the reference: the reference delegates to skimage.measure.marching_cubes_lewiner
(core/fusion.py:554-568, 579; requirements.txt:35 pins scikit-image 0.14), which is a third-party dependency absent from
/root/reference and from this image, and the reference holds no golden mesh.  What is kept from the call sites: the arguments
(`step_size`, `allow_degenerate=False`, no `level` -> skimage takes 0.5 * (min + max)), the 4-tuple result
(verts (V,3) float32 in voxel coordinates, faces (F,3) int32, normals (V,3) float32, values (V,) float32), an indexed mesh with
shared vertices (the reference averages face edge lengths, core/fusion.py:89-92, 592-596, and samples nodes from the vertices).
The mesh itself is DEFINED here (first principles, no table):

  * samples S = vol[::s, ::s, ::s]; a sample is ABOVE when S > level; an edge of the sampled grid that joins an ABOVE and a
    non-ABOVE sample owns one vertex at the linear crossing, t = (level - v0) / (v1 - v0) in float32, position (index + t) * s;
    vertices are ordered by owning (lower) sample in C order, then by axis;
  * normal = unit vector along the linearly interpolated gradient of S (central differences over the sampled grid, one-sided at
    the border, per voxel); value = the larger end of the edge;
  * inside a cell the crossed edges are joined face by face (a face with four crossings cuts off its ABOVE corners one by one),
    the resulting loops are taken in order of their smallest edge id (edge = 4 * axis + a + 2 * b), start there, are oriented along
    the +gradient and fan-triangulated; cells in C order; triangles with two vertices on the same grid sample (index + t rounds onto it) are dropped.
"""
import numpy as np

_OTHER = {0: (1, 2), 1: (0, 2), 2: (0, 1)}


def _edge_id(p, q):
    """Edge id of the cell edge between corner offsets p and q (tuples of 0/1 differing in one axis)."""
    d = [a for a in range(3) if p[a] != q[a]][0]
    a, b = _OTHER[d]
    return 4 * d + p[a] + 2 * p[b]


def _edge_ends(e):
    d, ab = divmod(e, 4)
    a, b = _OTHER[d]
    lo = [0, 0, 0]
    lo[a], lo[b] = ab & 1, ab >> 1
    hi = list(lo)
    hi[d] = 1
    return tuple(lo), tuple(hi)


def cell_loops(above):
    """above: dict corner offset (x,y,z) -> bool.  Oriented loops of edge ids."""
    link = {}
    for d in range(3):
        a, b = _OTHER[d]
        for side in (0, 1):
            ring = []
            for u, v in ((0, 0), (1, 0), (1, 1), (0, 1)):
                c = [0, 0, 0]
                c[d], c[a], c[b] = side, u, v
                ring.append(tuple(c))
            cut = [n for n in range(4) if above[ring[n]] != above[ring[(n + 1) % 4]]]
            name = [_edge_id(ring[n], ring[(n + 1) % 4]) for n in range(4)]
            if len(cut) == 2:
                segs = [(name[cut[0]], name[cut[1]])]
            elif len(cut) == 4:
                segs = [(name[n - 1], name[n]) for n in range(4) if above[ring[n]]]
            else:
                segs = []
            for p, q in segs:
                link.setdefault(p, []).append(q)
                link.setdefault(q, []).append(p)
    loops, used = [], set()
    for first in sorted(link):
        if first in used:
            continue
        loop = [first]
        while True:
            cand = [e for e in link[loop[-1]] if e != (loop[-2] if len(loop) > 1 else None)]
            nxt = cand[0]
            if nxt == first:
                break
            loop.append(nxt)
        used.update(loop)
        # the patch hangs inward from the cell face F holding the first segment: loop normal = F x T, to point BELOW -> ABOVE
        (a_lo, a_hi), (b_lo, b_hi) = _edge_ends(loop[0]), _edge_ends(loop[1])
        pts = np.array([a_lo, a_hi, b_lo, b_hi])
        F = np.zeros(3)
        for ax in range(3):
            if len(set(pts[:, ax])) == 1:
                F[ax] = 1.0 if pts[0, ax] else -1.0
        T = (pts[2] + pts[3]) / 2.0 - (pts[0] + pts[1]) / 2.0
        up = (np.array(a_lo) - np.array(a_hi)) if above[a_lo] else (np.array(a_hi) - np.array(a_lo))
        if np.cross(F, T) @ up < 0:
            loop = [loop[0]] + loop[1:][::-1]
        loops.append(loop)
    return loops


def case_table():
    """256 cases -> list of triangles (edge-id triples); case bit c = corner (c&1, (c>>1)&1, (c>>2)&1) above."""
    out = []
    for case in range(256):
        above = {(c & 1, (c >> 1) & 1, (c >> 2) & 1): bool((case >> c) & 1) for c in range(8)}
        tris = []
        for loop in cell_loops(above):
            tris += [(loop[0], loop[n], loop[n + 1]) for n in range(1, len(loop) - 1)]
        out.append(tris)
    return out


_TABLE = None


def _gradient(S, s):
    g = []
    for ax in range(3):
        n = S.shape[ax]
        out = np.zeros_like(S)
        if n >= 2:
            Sm = np.moveaxis(S, ax, 0)
            om = np.moveaxis(out, ax, 0)
            om[0] = (Sm[1] - Sm[0]) / np.float32(s)
            om[-1] = (Sm[-1] - Sm[-2]) / np.float32(s)
            if n >= 3:
                om[1:-1] = (Sm[2:] - Sm[:-2]) / np.float32(2 * s)
        g.append(out)
    return np.stack(g, axis=-1)


def default_level(vol):
    vol = np.asarray(vol, dtype=np.float32)
    return np.float32(0.5 * (np.float64(vol.min()) + np.float64(vol.max())))


def marching_cubes(vol, step_size=1, level=None, x_origin=0):
    """-> verts (V,3) f32, faces (F,3) i32, normals (V,3) f32, values (V,) f32.  x_origin: sample index of vol[0] in a larger grid
    (x-slabs): x coordinates, and the "index + t rounds onto a sample" test, use the index in that grid."""
    global _TABLE
    if _TABLE is None:
        _TABLE = case_table()
    vol = np.asarray(vol, dtype=np.float32)
    s = int(step_size)
    level = default_level(vol) if level is None else np.float32(level)
    S = np.ascontiguousarray(vol[::s, ::s, ::s])
    nx, ny, nz = S.shape
    above = S > level
    G = _gradient(S, s)
    lin = np.arange(S.size).reshape(S.shape)
    keys, pos, nrm, val = [], [], [], []
    for d in range(3):
        lo = [slice(None)] * 3
        hi = [slice(None)] * 3
        lo[d], hi[d] = slice(0, -1), slice(1, None)
        lo, hi = tuple(lo), tuple(hi)
        m = above[lo] != above[hi]
        idx = np.argwhere(m)
        v0, v1 = S[lo][m], S[hi][m]
        with np.errstate(all="ignore"):
            t = (level - v0) / (v1 - v0)
        idx = idx + np.array([int(x_origin), 0, 0])
        p = idx.astype(np.float32) * np.float32(s)
        p[:, d] = (idx[:, d].astype(np.float32) + t) * np.float32(s)
        g0, g1 = G[lo][m], G[hi][m]
        n = g0 + t[:, None] * (g1 - g0)
        ln = np.sqrt((n[:, 0] * n[:, 0] + n[:, 1] * n[:, 1]) + n[:, 2] * n[:, 2])
        with np.errstate(all="ignore"):
            n = np.where(ln[:, None] > 0, n / ln[:, None], np.float32(0))
        keys.append(lin[lo][m] * 3 + d)
        pos.append(p); nrm.append(n.astype(np.float32)); val.append(np.maximum(v0, v1))
    keys = np.concatenate(keys)
    order = np.argsort(keys, kind="stable")
    keys = keys[order]
    verts = np.concatenate(pos)[order].astype(np.float32).reshape(-1, 3)
    normals = np.concatenate(nrm)[order].reshape(-1, 3)
    values = np.concatenate(val)[order].astype(np.float32)

    faces = []
    if nx > 1 and ny > 1 and nz > 1:
        case = np.zeros((nx - 1, ny - 1, nz - 1), dtype=np.int32)
        for c in range(8):
            ox, oy, oz = c & 1, (c >> 1) & 1, (c >> 2) & 1
            case |= above[ox:nx - 1 + ox, oy:ny - 1 + oy, oz:nz - 1 + oz].astype(np.int32) << c
        for i, j, k in np.argwhere((case != 0) & (case != 255)):
            cs = int(case[i, j, k])
            for tri in _TABLE[cs]:
                ids, spots = [], []
                for e in tri:
                    lo, hi = _edge_ends(e)
                    d = e // 4
                    owner = (i + lo[0], j + lo[1], k + lo[2])
                    ids.append(int(np.searchsorted(keys, lin[owner] * 3 + d)))
                    far = (i + hi[0], j + hi[1], k + hi[2])
                    with np.errstate(all="ignore"):
                        xo = int(x_origin) if d == 0 else 0
                        u = np.float32(owner[d] + xo) + (level - S[owner]) / (S[far] - S[owner])
                    spots.append(owner if u == owner[d] + xo else far if u == far[d] + xo else ("edge", e))
                if len(set(spots)) == 3:
                    faces.append(ids)
    faces = np.asarray(faces, dtype=np.int32).reshape(-1, 3)
    return verts, faces, normals, values
