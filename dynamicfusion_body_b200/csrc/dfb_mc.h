// dfb_mc.h -- per-sample logic of the device surface extractor: the replacement of
// skimage.measure.marching_cubes_lewiner(volume, step_size=s, allow_degenerate=False) as the reference calls it
// (core/fusion.py:554-568, 579; level = mean of the volume's min and max because the reference passes none).
//
// Indexed mesh without a welding pass: every edge of the sampled grid that crosses the level owns one vertex, and the edge belongs
// to its lower sample (i,j,k) -- so a vertex id is a pure function of (sample, axis) and a prefix count:
//     id = row_voff[row(i,j)] + chunk.voff + (crossed edges of the row's samples before k in the 32-sample chunk) + (axes < d at k).
// Vertex order: sample-major (x, y, z), axis-minor.  Face order: cell-major, then the order of dfb_mc_table.h.
// Triangles with two vertices on the same grid sample (a corner value at, or within rounding of, the level) are dropped
// (allow_degenerate=False).
//
// Like dfb_math.h this header also compiles for the host (tests/hostshim/) so the CPU suite runs the same functions.
#pragma once
#include "dfb_math.h"
#include "dfb_mc_table.h"

namespace dfb {

struct McGrid {
    const float* vol;  // [rx][ry][rz], z fastest
    int rx, ry, rz;
    int step;          // sampling stride (skimage step_size)
    int nx, ny, nz;    // samples per axis: (r - 1) / step + 1; cells per axis: n - 1
    int ncz;           // 32-sample chunks per row: ceil(nz / 32)
    float level;
    int xs0;           // sample index of the volume's first x-plane in the full grid (x-slabs: vertex coordinates and the
                       // degenerate-triangle test are those of the whole grid); 0 for a whole volume
};

struct McChunk {       // one per 32 consecutive samples of a row
    int32_t voff;      // vertices of the row before this chunk
    uint32_t m[3];     // bit l: the edge along axis d owned by sample 32*c + l crosses the level
};

#if defined(__CUDACC__)
__device__ const uint8_t d_mc_ntri[256] = DFB_MC_NTRI_INIT;
__device__ const int8_t d_mc_tri[256 * 3 * DFB_MC_MAX_TRIS] = DFB_MC_TRI_INIT;
#endif
static const uint8_t h_mc_ntri[256] = DFB_MC_NTRI_INIT;
static const int8_t h_mc_tri[256 * 3 * DFB_MC_MAX_TRIS] = DFB_MC_TRI_INIT;

DFB_HD int mc_table_ntri(int c) {
#if defined(__CUDA_ARCH__)
    return d_mc_ntri[c];
#else
    return h_mc_ntri[c];
#endif
}
DFB_HD int mc_table_edge(int c, int t) {
#if defined(__CUDA_ARCH__)
    return d_mc_tri[c * (3 * DFB_MC_MAX_TRIS) + t];
#else
    return h_mc_tri[c * (3 * DFB_MC_MAX_TRIS) + t];
#endif
}
DFB_HD int mc_popc(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}

DFB_HD void mc_grid_init(McGrid& g, const float* vol, int rx, int ry, int rz, int step, float level, int xs0 = 0) {
    g.vol = vol; g.rx = rx; g.ry = ry; g.rz = rz; g.step = step; g.level = level; g.xs0 = xs0;
    g.nx = (rx - 1) / step + 1; g.ny = (ry - 1) / step + 1; g.nz = (rz - 1) / step + 1;
    g.ncz = (g.nz + 31) / 32;
}

DFB_HD float mc_val(const McGrid& g, int i, int j, int k) {
    return g.vol[((size_t)(i * g.step) * g.ry + (size_t)(j * g.step)) * g.rz + (size_t)(k * g.step)];
}

// bit d: the edge from sample (i,j,k) to its +d neighbour exists and crosses the level
DFB_HD uint32_t mc_edge_bits(const McGrid& g, int i, int j, int k, float v0) {
    const bool a0 = v0 > g.level;
    uint32_t bits = 0;
    if (i + 1 < g.nx && (mc_val(g, i + 1, j, k) > g.level) != a0) bits |= 1u;
    if (j + 1 < g.ny && (mc_val(g, i, j + 1, k) > g.level) != a0) bits |= 2u;
    if (k + 1 < g.nz && (mc_val(g, i, j, k + 1) > g.level) != a0) bits |= 4u;
    return bits;
}

// the 8 corner values of cell (i,j,k) (corner c at offset (c&1, (c>>1)&1, (c>>2)&1)) and its case
DFB_HD int mc_cell_case(const McGrid& g, int i, int j, int k, float v[8]) {
    int cs = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        v[c] = mc_val(g, i + (c & 1), j + ((c >> 1) & 1), k + ((c >> 2) & 1));
        cs |= (v[c] > g.level ? 1 : 0) << c;
    }
    return cs;
}

// corner offsets of edge e = 4*d + a + 2*b: lower corner; the upper one adds (1 << d)
DFB_HD int mc_edge_lower_corner(int e) {
    const int d = e >> 2, a = e & 1, b = (e >> 1) & 1;
    return d == 0 ? ((a << 1) | (b << 2)) : d == 1 ? (a | (b << 2)) : (a | (b << 1));
}

// coordinate of a crossing along its edge: (index of the lower sample + t) with t = (level - v0) / (v1 - v0), float32
DFB_HD float mc_cross_coord(int base, float v0, float v1, float level) {
    return fadd((float)base, fdiv(fsub(level, v0), fsub(v1, v0)));
}

// key of the position an edge vertex takes: the corner it coincides with when the rounded crossing coordinate lands on a grid
// sample (a corner value at, or within rounding of, the level), else a value unique to the edge.  ijk: the cell.
DFB_HD int mc_edge_key(int e, const float v[8], float level, const int ijk[3]) {
    const int d = e >> 2, c0 = mc_edge_lower_corner(e), c1 = c0 | (1 << d);
    const float u = mc_cross_coord(ijk[d], v[c0], v[c1], level);
    return u == (float)ijk[d] ? c0 : u == (float)(ijk[d] + 1) ? c1 : 8 + e;
}

// non-degenerate triangles of cell (i,j,k) (i counted in the full grid: local index + McGrid::xs0): writes 3 edge ids per triangle, returns their number
DFB_HD int mc_cell_tris(int cs, const float v[8], float level, int i, int j, int k, int8_t* edges /* [3 * DFB_MC_MAX_TRIS] or null */) {
    const int n = mc_table_ntri(cs);
    const int ijk[3] = {i, j, k};
    int kept = 0;
    for (int t = 0; t < n; ++t) {
        const int e0 = mc_table_edge(cs, 3 * t), e1 = mc_table_edge(cs, 3 * t + 1), e2 = mc_table_edge(cs, 3 * t + 2);
        const int k0 = mc_edge_key(e0, v, level, ijk), k1 = mc_edge_key(e1, v, level, ijk), k2 = mc_edge_key(e2, v, level, ijk);
        if (k0 == k1 || k0 == k2 || k1 == k2) continue;
        if (edges) { edges[3 * kept] = (int8_t)e0; edges[3 * kept + 1] = (int8_t)e1; edges[3 * kept + 2] = (int8_t)e2; }
        ++kept;
    }
    return kept;
}

DFB_HD int mc_vertex_id(const McGrid& g, const McChunk* chunks, const int32_t* row_voff, int i, int j, int k, int d) {
    const int row = i * g.ny + j;
    const McChunk rec = chunks[(size_t)row * g.ncz + (k >> 5)];
    const int lane = k & 31;
    const uint32_t lt = (1u << lane) - 1u;
    int id = row_voff[row] + rec.voff + mc_popc(rec.m[0] & lt) + mc_popc(rec.m[1] & lt) + mc_popc(rec.m[2] & lt);
    for (int dd = 0; dd < d; ++dd) id += (rec.m[dd] >> lane) & 1u;
    return id;
}

// vertex id of edge e of cell (i,j,k)
DFB_HD int mc_cell_edge_vertex(const McGrid& g, const McChunk* chunks, const int32_t* row_voff, int i, int j, int k, int e) {
    const int c0 = mc_edge_lower_corner(e);
    return mc_vertex_id(g, chunks, row_voff, i + (c0 & 1), j + ((c0 >> 1) & 1), k + ((c0 >> 2) & 1), e >> 2);
}

// central difference over the sampled grid (one-sided at the border), in value units per voxel
DFB_HD float mc_diff(float vm, float v0, float vp, bool has_m, bool has_p, int step) {
    if (has_m && has_p) return fdiv(fsub(vp, vm), (float)(2 * step));
    if (has_p) return fdiv(fsub(vp, v0), (float)step);
    if (has_m) return fdiv(fsub(v0, vm), (float)step);
    return 0.0f;
}

DFB_HD void mc_gradient(const McGrid& g, int i, int j, int k, float out[3]) {
    const float v0 = mc_val(g, i, j, k);
    out[0] = mc_diff(i > 0 ? mc_val(g, i - 1, j, k) : 0.f, v0, i + 1 < g.nx ? mc_val(g, i + 1, j, k) : 0.f, i > 0, i + 1 < g.nx, g.step);
    out[1] = mc_diff(j > 0 ? mc_val(g, i, j - 1, k) : 0.f, v0, j + 1 < g.ny ? mc_val(g, i, j + 1, k) : 0.f, j > 0, j + 1 < g.ny, g.step);
    out[2] = mc_diff(k > 0 ? mc_val(g, i, j, k - 1) : 0.f, v0, k + 1 < g.nz ? mc_val(g, i, j, k + 1) : 0.f, k > 0, k + 1 < g.nz, g.step);
}

// the vertex of the crossed edge along axis d owned by sample (i,j,k): position in voxel coordinates, unit normal along the
// interpolated +gradient (zero when it vanishes), value = the larger end
DFB_HD void mc_vertex(const McGrid& g, int i, int j, int k, int d, float pos[3], float nrm[3], float& value) {
    const int i1 = i + (d == 0), j1 = j + (d == 1), k1 = k + (d == 2);
    const float v0 = mc_val(g, i, j, k), v1 = mc_val(g, i1, j1, k1);
    const float t = fdiv(fsub(g.level, v0), fsub(v1, v0));
    const float fs = (float)g.step;
    pos[0] = fmul((float)(i + g.xs0), fs); pos[1] = fmul((float)j, fs); pos[2] = fmul((float)k, fs);
    pos[d] = fmul(mc_cross_coord(d == 0 ? i + g.xs0 : d == 1 ? j : k, v0, v1, g.level), fs);
    float g0[3], g1[3], n[3];
    mc_gradient(g, i, j, k, g0);
    mc_gradient(g, i1, j1, k1, g1);
    for (int a = 0; a < 3; ++a) n[a] = fadd(g0[a], fmul(t, fsub(g1[a], g0[a])));
    const float len = fsqrt(fadd(fadd(fmul(n[0], n[0]), fmul(n[1], n[1])), fmul(n[2], n[2])));
    for (int a = 0; a < 3; ++a) nrm[a] = len > 0.0f ? fdiv(n[a], len) : 0.0f;
    value = v0 > v1 ? v0 : v1;
}

}  // namespace dfb
