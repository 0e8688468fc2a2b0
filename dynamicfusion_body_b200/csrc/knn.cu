// knn.cu -- exact k-nearest deformation nodes (scipy KDTree.query semantics, core/fusion.py:175,122).
//
// The reference queries a KD-tree per voxel; the result is a pure function of the node positions, which
// only change in update_graph (core/fusion.py:201-233), so the table is built once per graph revision
// and reused by every frame / view (uint16 ids, 2k bytes per voxel).
//
// Exactness: distances are ranked in fp32 first (node tiles staged in shared memory, one broadcast
// LDS.128 per node per warp); whenever two of the k+1 best squared distances are closer than the fp32
// rounding bound the voxel is re-ranked in float64 with the oracle's operation order and the "lower id
// wins" tie rule.
#include <stdlib.h>

#include "common.h"
#include "dfb_math.h"

using namespace dfb;

namespace {

constexpr int TILE = 1024;

template <int KMAX>
struct TopK {
    float d[KMAX + 1];
    int id[KMAX + 1];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int j = 0; j <= KMAX; ++j) { d[j] = 3.0e38f; id[j] = 0; }
    }
    // keep the kk+1 smallest (kk <= KMAX), ascending, stable in id order (strict <)
    __device__ __forceinline__ void insert(float v, int i, int kk) {
        if (!(v < d[kk])) return;
        d[kk] = v; id[kk] = i;
#pragma unroll
        for (int j = KMAX; j > 0; --j) {
            if (j <= kk && d[j] < d[j - 1]) {
                const float td = d[j]; d[j] = d[j - 1]; d[j - 1] = td;
                const int ti = id[j]; id[j] = id[j - 1]; id[j - 1] = ti;
            }
        }
    }
};

// float64 re-rank of one query (rare path).  out: k ids ascending, ties -> lower id.
template <int KMAX>
__device__ __noinline__ void knn_exact_f64(float qx, float qy, float qz, const float* node_pos, int n, int k, int* out) {
    double bd[KMAX];
    int bi[KMAX];
    for (int j = 0; j < KMAX; ++j) { bd[j] = 1.0e300; bi[j] = 0; }
    for (int i = 0; i < n; ++i) {
        const double dx = dfb::dsub((double)qx, (double)node_pos[3 * i]);
        const double dy = dfb::dsub((double)qy, (double)node_pos[3 * i + 1]);
        const double dz = dfb::dsub((double)qz, (double)node_pos[3 * i + 2]);
        const double d2 = dfb::dadd(dfb::dadd(dfb::dmul(dx, dx), dfb::dmul(dy, dy)), dfb::dmul(dz, dz));
        if (!(d2 < bd[k - 1])) continue;
        int j = k - 1;
        while (j > 0 && d2 < bd[j - 1]) { bd[j] = bd[j - 1]; bi[j] = bi[j - 1]; --j; }
        bd[j] = d2; bi[j] = i;
    }
    for (int j = 0; j < k; ++j) out[j] = bi[j];
}

template <int KMAX>
__device__ __forceinline__ void knn_query(float qx, float qy, float qz, bool active, const float* node_pos, int n, int k,
                                          float4* tile, int* out) {
    TopK<KMAX> top;
    top.init();
    const int kk = (n > k) ? k : k - 1;  // index of the last tracked slot (k+1 entries when available)
    for (int base = 0; base < n; base += TILE) {
        const int cnt = min(TILE, n - base);
        __syncthreads();
        for (int j = threadIdx.x; j < cnt; j += blockDim.x)
            tile[j] = make_float4(node_pos[3 * (base + j)], node_pos[3 * (base + j) + 1], node_pos[3 * (base + j) + 2], 0.f);
        __syncthreads();
        if (active) {
#pragma unroll 4
            for (int j = 0; j < cnt; ++j) {
                const float4 p = tile[j];
                const float dx = qx - p.x, dy = qy - p.y, dz = qz - p.z;
                const float d2 = dx * dx + dy * dy + dz * dz;
                top.insert(d2, base + j, kk);
            }
        }
    }
    if (!active) return;
    bool ambiguous = false;
#pragma unroll
    for (int j = 0; j < KMAX; ++j)
        if (j < kk && top.d[j + 1] - top.d[j] <= 4.0e-6f * top.d[j + 1]) ambiguous = true;
    if (ambiguous) {
        knn_exact_f64<KMAX>(qx, qy, qz, node_pos, n, k, out);
    } else {
#pragma unroll
        for (int j = 0; j < KMAX; ++j)
            if (j < k) out[j] = top.id[j];
    }
}

template <int KMAX>
__global__ void __launch_bounds__(256) knn_volume_kernel(const float* node_pos, int n, int k, int ry, int rz, int x0,
                                                         uint16_t* knn) {
    __shared__ float4 tile[TILE];
    const int z = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int xs = blockIdx.z;
    const bool in = z < rz;
    int out[KMAX];
    knn_query<KMAX>((float)(xs + x0), (float)y, (float)z, in, node_pos, n, k, tile, out);
    if (!in) return;
    const size_t i = ((size_t)xs * ry + y) * rz + z;
    if (KMAX == 4 && k == 4) {
        uint2 r;
        r.x = (uint32_t)out[0] | ((uint32_t)out[1] << 16);
        r.y = (uint32_t)out[2] | ((uint32_t)out[3] << 16);
        reinterpret_cast<uint2*>(knn)[i] = r;
    } else {
#pragma unroll
        for (int j = 0; j < KMAX; ++j)
            if (j < k) knn[i * (size_t)k + j] = (uint16_t)out[j];
    }
}


// ------------------------------------------------------------------------------------------------------------------
// Brick-accelerated exact build.  One CTA per 8x8x8 brick: (1) d_k of the brick centre c by a block-wide selection,
// (2) candidate nodes = { n : |n - c| <= d_k(c) + 2 * halfdiag } staged in shared memory -- every node among the k
// nearest of ANY voxel of the brick is in there (d_k(v) <= d_k(c) + h and |n - c| <= |n - v| + h) -- (3) per-voxel
// ranking over the candidates only, with the same fp32-then-float64 tie discipline as the brute-force kernel.
// ------------------------------------------------------------------------------------------------------------------
constexpr int KB = 8;            // brick edge
constexpr int KCAP = 1024;       // candidate capacity (16 KB); bricks that overflow fall back to the full node list

template <int KMAX>
__device__ __noinline__ void knn_exact_f64_list(float qx, float qy, float qz, const float4* cand, int nc, int k, int* out) {
    double bd[KMAX];
    int bi[KMAX];
    for (int j = 0; j < KMAX; ++j) { bd[j] = 1.0e300; bi[j] = 0x7fffffff; }
    for (int t = 0; t < nc; ++t) {
        const float4 p = cand[t];
        const int id = __float_as_int(p.w);
        const double dx = dfb::dsub((double)qx, (double)p.x), dy = dfb::dsub((double)qy, (double)p.y), dz = dfb::dsub((double)qz, (double)p.z);
        const double d2 = dfb::dadd(dfb::dadd(dfb::dmul(dx, dx), dfb::dmul(dy, dy)), dfb::dmul(dz, dz));
        if (!(d2 < bd[k - 1] || (d2 == bd[k - 1] && id < bi[k - 1]))) continue;
        int j = k - 1;
        while (j > 0 && (d2 < bd[j - 1] || (d2 == bd[j - 1] && id < bi[j - 1]))) { bd[j] = bd[j - 1]; bi[j] = bi[j - 1]; --j; }
        bd[j] = d2; bi[j] = id;
    }
    for (int j = 0; j < k; ++j) out[j] = bi[j];
}

// brick_T (optional, out): the search radius T of every brick, kept for incremental graph revisions (knn_merge_kernel).  With n_old > 0
// the kernel is the second pass of such a revision: it rebuilds exactly the bricks the merge pass flagged with dirty == 2.
template <int KMAX>
__global__ void __launch_bounds__(128) knn_brick_kernel(const float* node_pos, int n, int k, int sx, int ry, int rz, int x0, int nby, int nbz,
                                                        uint16_t* knn, float* brick_T, int n_old, uint8_t* dirty) {
    __shared__ float4 cand[KCAP];
    __shared__ float wtop[4][KMAX];
    __shared__ int ncand;
    __shared__ float T2s;
    const int b = blockIdx.x;
    const int bz = b % nbz, by = (b / nbz) % nby, bxs = b / (nbz * nby);
    const int xlo = bxs * KB, ylo = by * KB, zlo = bz * KB;
    const int xhi = min(xlo + KB, sx) - 1, yhi = min(ylo + KB, ry) - 1, zhi = min(zlo + KB, rz) - 1;
    const float cx = 0.5f * (xlo + xhi) + (float)x0, cy = 0.5f * (ylo + yhi), cz = 0.5f * (zlo + zhi);
    const float hx = 0.5f * (xhi - xlo), hy = 0.5f * (yhi - ylo), hz = 0.5f * (zhi - zlo);
    const float hd = sqrtf(hx * hx + hy * hy + hz * hz);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (n_old > 0) {
        // incremental pass 2: only the bricks knn_merge_kernel could not handle (more new nodes in reach than its list holds)
        if (dirty[b] != 2) return;
        __syncthreads();
        if (threadIdx.x == 0) dirty[b] = 1;
    } else if (threadIdx.x == 0 && dirty) {
        dirty[b] = 1;
    }
    // (1) k smallest squared distances to the centre: per-thread sorted list, warp merge, then 4-way merge
    float lt[KMAX];
#pragma unroll
    for (int j = 0; j < KMAX; ++j) lt[j] = 3.0e38f;
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        const float dx = node_pos[3 * t] - cx, dy = node_pos[3 * t + 1] - cy, dz = node_pos[3 * t + 2] - cz;
        float v = dx * dx + dy * dy + dz * dz;
#pragma unroll
        for (int j = 0; j < KMAX; ++j)
            if (j < k && v < lt[j]) { const float tmp = lt[j]; lt[j] = v; v = tmp; }
    }
    for (int j = 0; j < k; ++j) {
        float m = lt[0];
        for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
        const unsigned who = __ballot_sync(0xffffffffu, lt[0] == m);
        if (lane == __ffs(who) - 1) {   // pop the winner's head
#pragma unroll
            for (int q = 0; q < KMAX - 1; ++q) lt[q] = lt[q + 1];
            lt[KMAX - 1] = 3.0e38f;
        }
        if (lane == 0) wtop[wid][j] = m;
    }
    if (threadIdx.x == 0) ncand = 0;
    __syncthreads();
    if (threadIdx.x == 0) {
        float all[4 * KMAX];
        int m = 0;
        for (int w = 0; w < 4; ++w)
            for (int j = 0; j < k; ++j) all[m++] = wtop[w][j];
        for (int a = 1; a < m; ++a) {   // insertion sort of <= 32 values
            const float v = all[a];
            int q = a - 1;
            while (q >= 0 && all[q] > v) { all[q + 1] = all[q]; --q; }
            all[q + 1] = v;
        }
        const float T = sqrtf(all[k - 1]) * 1.00001f + 2.f * hd + 1e-3f;
        T2s = T * T * 1.00001f;
        if (brick_T) brick_T[b] = T;
    }
    __syncthreads();
    const float T2 = T2s;
    // (2) candidates
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        const float px = node_pos[3 * t], py = node_pos[3 * t + 1], pz = node_pos[3 * t + 2];
        const float dx = px - cx, dy = py - cy, dz = pz - cz;
        if (dx * dx + dy * dy + dz * dz <= T2) {
            const int pos = atomicAdd(&ncand, 1);
            if (pos < KCAP) cand[pos] = make_float4(px, py, pz, __int_as_float(t));
        }
    }
    __syncthreads();
    const int nc = ncand;
    const bool overflow = nc > KCAP;   // block-uniform
    // (3) per-voxel ranking: thread t owns voxels j = 4t .. 4t+3 of the brick (z fastest)
    TopK<KMAX> top[4];
    float qx[4], qy[4], qz[4];
    bool in[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int j = threadIdx.x * 4 + q;
        const int z = zlo + (j & 7), y = ylo + ((j >> 3) & 7), xs = xlo + (j >> 6);
        in[q] = xs < sx && y < ry && z < rz;
        qx[q] = (float)(xs + x0); qy[q] = (float)y; qz[q] = (float)z;
        top[q].init();
    }
    const int ntot = overflow ? n : nc;
    const int kk = (ntot > k) ? k : k - 1;
    if (!overflow) {
        for (int t = 0; t < nc; ++t) {
            const float4 p = cand[t];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float dx = qx[q] - p.x, dy = qy[q] - p.y, dz = qz[q] - p.z;
                top[q].insert(dx * dx + dy * dy + dz * dz, __float_as_int(p.w), kk);
            }
        }
    } else {
        // too many candidates for shared memory (bricks far from every node): stream the whole node list through it
        for (int base = 0; base < n; base += KCAP) {
            const int cnt = min(KCAP, n - base);
            __syncthreads();
            for (int t = threadIdx.x; t < cnt; t += blockDim.x)
                cand[t] = make_float4(node_pos[3 * (base + t)], node_pos[3 * (base + t) + 1], node_pos[3 * (base + t) + 2], __int_as_float(base + t));
            __syncthreads();
            for (int t = 0; t < cnt; ++t) {
                const float4 p = cand[t];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float dx = qx[q] - p.x, dy = qy[q] - p.y, dz = qz[q] - p.z;
                    top[q].insert(dx * dx + dy * dy + dz * dz, base + t, kk);
                }
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        if (!in[q]) continue;
        int out[KMAX];
        bool ambiguous = false;
#pragma unroll
        for (int a = 0; a < KMAX; ++a)
            if (a < kk && top[q].d[a + 1] - top[q].d[a] <= 4.0e-6f * top[q].d[a + 1]) ambiguous = true;
        if (ambiguous) {
            if (overflow) knn_exact_f64<KMAX>(qx[q], qy[q], qz[q], node_pos, n, k, out);
            else knn_exact_f64_list<KMAX>(qx[q], qy[q], qz[q], cand, nc, k, out);
        } else {
#pragma unroll
            for (int a = 0; a < KMAX; ++a)
                if (a < k) out[a] = top[q].id[a];
        }
        const int j = threadIdx.x * 4 + q;
        const int z = zlo + (j & 7), y = ylo + ((j >> 3) & 7), xs = xlo + (j >> 6);
        const size_t i = ((size_t)xs * ry + y) * rz + z;
        if (KMAX == 4 && k == 4) {
            uint2 r;
            r.x = (uint32_t)out[0] | ((uint32_t)out[1] << 16);
            r.y = (uint32_t)out[2] | ((uint32_t)out[3] << 16);
            reinterpret_cast<uint2*>(knn)[i] = r;
        } else {
#pragma unroll
            for (int a = 0; a < KMAX; ++a)
                if (a < k) knn[i * (size_t)k + a] = (uint16_t)out[a];
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Incremental graph revision (core/fusion.py:216-229 only ever APPENDS nodes).  The table holds the k nearest of the first n_old
// nodes, brick_T the search radius T = d_k(centre) + 2 halfdiag every brick was built with.  A node appended since can be among the k
// nearest of some voxel of the brick only if it lies within T of the centre (it would have been a candidate); a brick with no such
// node leaves at once -- its rows are still exact: old ids are unchanged, and on an exact distance tie the lower, i.e. old, id wins.
// Otherwise every voxel merges the few new nodes in reach into its stored, already correctly ordered list: float64 squared distances
// in the oracle's operation order, strict '<' against the old entries (ties keep the old node in front), new nodes taken in
// ascending id.  No scan over the node table, no candidate search: the pass is bounded by the 8 bytes per voxel it reads and, where
// a row changed, writes.  dirty[b] = 1 when a row of the brick changed (the 4x4x32 brick / region sets that derive from the table are
// refreshed for those), 2 when more than KNEW new nodes are in reach (knn_brick_kernel rebuilds such a brick from scratch).
// ------------------------------------------------------------------------------------------------------------------
constexpr int KNEW = 48;
template <int KMAX>
__global__ void __launch_bounds__(128) knn_merge_kernel(const float* node_pos, int n_old, int n, int k, int sx, int ry, int rz, int x0, int nby, int nbz,
                                                        uint16_t* knn, const float* brick_T, uint8_t* dirty) {
    __shared__ float4 newc[KNEW];
    __shared__ int n_new_s, changed_s;
    const int b = blockIdx.x;
    const int bz = b % nbz, by = (b / nbz) % nby, bxs = b / (nbz * nby);
    const int xlo = bxs * KB, ylo = by * KB, zlo = bz * KB;
    const int xhi = min(xlo + KB, sx) - 1, yhi = min(ylo + KB, ry) - 1, zhi = min(zlo + KB, rz) - 1;
    const float cx = 0.5f * (xlo + xhi) + (float)x0, cy = 0.5f * (ylo + yhi), cz = 0.5f * (zlo + zhi);
    if (threadIdx.x == 0) { n_new_s = 0; changed_s = 0; }
    __syncthreads();
    const float T = brick_T[b];
    const float T2 = T * T * 1.00001f;
    for (int t = n_old + threadIdx.x; t < n; t += blockDim.x) {
        const float px = node_pos[3 * t], py = node_pos[3 * t + 1], pz = node_pos[3 * t + 2];
        const float dx = px - cx, dy = py - cy, dz = pz - cz;
        if (dx * dx + dy * dy + dz * dz <= T2) {
            const int pos = atomicAdd(&n_new_s, 1);
            if (pos < KNEW) newc[pos] = make_float4(px, py, pz, __int_as_float(t));
        }
    }
    __syncthreads();
    const int m = n_new_s;
    if (m == 0 || m > KNEW) {
        if (threadIdx.x == 0) dirty[b] = m == 0 ? 0 : 2;
        return;
    }
    if (threadIdx.x == 0) {   // ascending id (the appends raced): insertion sort of <= KNEW entries
        for (int a = 1; a < m; ++a) {
            const float4 v = newc[a];
            int q = a - 1;
            while (q >= 0 && __float_as_int(newc[q].w) > __float_as_int(v.w)) { newc[q + 1] = newc[q]; --q; }
            newc[q + 1] = v;
        }
    }
    __syncthreads();
    bool changed = false;
    for (int q = 0; q < 4; ++q) {
        const int j = threadIdx.x * 4 + q;
        const int z = zlo + (j & 7), y = ylo + ((j >> 3) & 7), xs = xlo + (j >> 6);
        if (!(xs < sx && y < ry && z < rz)) continue;
        const size_t i = ((size_t)xs * ry + y) * rz + z;
        const double qx = (double)(float)(xs + x0), qy = (double)y, qz = (double)z;
        int id[KMAX];
        double d[KMAX];
#pragma unroll
        for (int a = 0; a < KMAX; ++a) {
            id[a] = 0; d[a] = 1.0e300;
            if (a < k) {
                id[a] = knn[i * (size_t)k + a];
                const float* np_ = node_pos + 3 * (size_t)id[a];
                const double dx = dfb::dsub(qx, (double)np_[0]), dy = dfb::dsub(qy, (double)np_[1]), dz = dfb::dsub(qz, (double)np_[2]);
                d[a] = dfb::dadd(dfb::dadd(dfb::dmul(dx, dx), dfb::dmul(dy, dy)), dfb::dmul(dz, dz));
            }
        }
        bool ch = false;
        for (int t = 0; t < m; ++t) {
            const float4 p = newc[t];
            const double dx = dfb::dsub(qx, (double)p.x), dy = dfb::dsub(qy, (double)p.y), dz = dfb::dsub(qz, (double)p.z);
            const double dn = dfb::dadd(dfb::dadd(dfb::dmul(dx, dx), dfb::dmul(dy, dy)), dfb::dmul(dz, dz));
            if (!(dn < d[k - 1])) continue;          // not strictly closer than the current k-th: stays out (ties keep the lower id)
            // insert before the first entry that is strictly farther; entries at the same distance (lower ids) stay in front
            int idn = __float_as_int(p.w);
            double dcur = dn;
            bool ins = false;
#pragma unroll
            for (int a = 0; a < KMAX; ++a) {
                if (a < k) {
                    if (!ins && dcur < d[a]) ins = true;
                    if (ins) {                       // from the insertion point on everything moves down one slot (order of ties kept)
                        const double td = d[a]; d[a] = dcur; dcur = td;
                        const int ti = id[a]; id[a] = idn; idn = ti;
                    }
                }
            }
            ch = true;
        }
        if (ch) {
#pragma unroll
            for (int a = 0; a < KMAX; ++a)
                if (a < k) knn[i * (size_t)k + a] = (uint16_t)id[a];
            changed = true;
        }
    }
    if (changed) changed_s = 1;
    __syncthreads();
    if (threadIdx.x == 0) dirty[b] = changed_s ? 1 : 0;
}

template <int KMAX>
__global__ void __launch_bounds__(256) knn_points_kernel(const float* pts, int64_t m, const float* node_pos, int n, int k,
                                                         int32_t* idx) {
    __shared__ float4 tile[TILE];
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = t < m;
    const int64_t tt = in ? t : 0;
    int out[KMAX];
    knn_query<KMAX>(pts[3 * tt], pts[3 * tt + 1], pts[3 * tt + 2], in, node_pos, n, k, tile, out);
    if (!in) return;
#pragma unroll
    for (int j = 0; j < KMAX; ++j)
        if (j < k) idx[t * k + j] = out[j];
}

}  // namespace

namespace {
int knn_build(const float* node_pos, int n_old, int n_nodes, int k, int rx, int ry, int rz, int x0, int x1, uint16_t* knn, float* brick_T,
              uint8_t* dirty, dfb_stream_t stream) {
    DFB_REQUIRE(node_pos && knn, "null pointer");
    DFB_REQUIRE(k >= 1 && k <= DFB_MAX_K, "k=%d out of range [1,%d]", k, DFB_MAX_K);
    DFB_REQUIRE(n_nodes >= k && n_nodes <= 65535, "n_nodes=%d must be in [k,65535]", n_nodes);
    DFB_REQUIRE(rx > 0 && ry > 0 && rz > 0 && x0 >= 0 && x1 > x0 && x1 <= rx, "bad grid / slab");
    DFB_REQUIRE(ry <= 65535 && (x1 - x0) <= 65535, "ry and slab thickness must be <= 65535");
    DFB_REQUIRE(n_old == 0 || (n_old >= k && n_old <= n_nodes && brick_T), "incremental build needs the radii of a table over >= k nodes");
    cudaStream_t s = (cudaStream_t)stream;
    // DFB_KNN_BRUTE=1 selects the O(voxels * nodes) reference kernel (validation of the brick build)
    const char* env = getenv("DFB_KNN_BRUTE");
    const int brute = (env && atoi(env)) ? 1 : 0;
    const int64_t nbx = (x1 - x0 + KB - 1) / KB, nby = (ry + KB - 1) / KB, nbz = (rz + KB - 1) / KB;
    if ((brute && n_old == 0 && !brick_T) || nbx * nby * nbz >= ((int64_t)1 << 31)) {
        DFB_REQUIRE(n_old == 0 && !brick_T, "slab too large for the brick build");
        const int threads = rz >= 256 ? 256 : ((rz + 31) / 32) * 32;
        const dim3 grid((rz + threads - 1) / threads, ry, x1 - x0);
        if (k <= 4) knn_volume_kernel<4><<<grid, threads, 0, s>>>(node_pos, n_nodes, k, ry, rz, x0, knn);
        else knn_volume_kernel<8><<<grid, threads, 0, s>>>(node_pos, n_nodes, k, ry, rz, x0, knn);
        DFB_LAUNCH_CHECK("knn_volume_kernel");
        return DFB_OK;
    }
    const unsigned nb = (unsigned)(nbx * nby * nbz);
    if (n_old > 0) {
        if (k <= 4) knn_merge_kernel<4><<<nb, 128, 0, s>>>(node_pos, n_old, n_nodes, k, x1 - x0, ry, rz, x0, (int)nby, (int)nbz, knn, brick_T, dirty);
        else knn_merge_kernel<8><<<nb, 128, 0, s>>>(node_pos, n_old, n_nodes, k, x1 - x0, ry, rz, x0, (int)nby, (int)nbz, knn, brick_T, dirty);
        DFB_LAUNCH_CHECK("knn_merge_kernel");
        if (n_nodes - n_old <= KNEW) return DFB_OK;      // no brick can have overflowed: the rebuild pass would find nothing to do
    }
    if (k <= 4) knn_brick_kernel<4><<<nb, 128, 0, s>>>(node_pos, n_nodes, k, x1 - x0, ry, rz, x0, (int)nby, (int)nbz, knn, brick_T, n_old, dirty);
    else knn_brick_kernel<8><<<nb, 128, 0, s>>>(node_pos, n_nodes, k, x1 - x0, ry, rz, x0, (int)nby, (int)nbz, knn, brick_T, n_old, dirty);
    DFB_LAUNCH_CHECK("knn_brick_kernel");
    return DFB_OK;
}
}  // namespace

extern "C" int dfb_knn_build_volume(const float* node_pos, int n_nodes, int k, int rx, int ry, int rz, int x0, int x1,
                                    uint16_t* knn, dfb_stream_t stream) {
    return knn_build(node_pos, 0, n_nodes, k, rx, ry, rz, x0, x1, knn, nullptr, nullptr, stream);
}

extern "C" int64_t dfb_knn_brick_count(int slab_x, int ry, int rz) {
    return (int64_t)((slab_x + KB - 1) / KB) * ((ry + KB - 1) / KB) * ((rz + KB - 1) / KB);
}

extern "C" int dfb_knn_build_volume_radii(const float* node_pos, int n_nodes, int k, int rx, int ry, int rz, int x0, int x1, uint16_t* knn,
                                          float* brick_radius, dfb_stream_t stream) {
    DFB_REQUIRE(brick_radius, "null pointer");
    return knn_build(node_pos, 0, n_nodes, k, rx, ry, rz, x0, x1, knn, brick_radius, nullptr, stream);
}

extern "C" int dfb_knn_update_volume(const float* node_pos, int n_old, int n_nodes, int k, int rx, int ry, int rz, int x0, int x1, uint16_t* knn,
                                     float* brick_radius, uint8_t* dirty, dfb_stream_t stream) {
    DFB_REQUIRE(brick_radius && dirty, "null pointer");
    DFB_REQUIRE(n_old > 0, "n_old must be positive (use dfb_knn_build_volume_radii for the first build)");
    return knn_build(node_pos, n_old, n_nodes, k, rx, ry, rz, x0, x1, knn, brick_radius, dirty, stream);
}

extern "C" int dfb_knn_points(const float* pts, int64_t m, const float* node_pos, int n_nodes, int k, int32_t* idx,
                              dfb_stream_t stream) {
    DFB_REQUIRE(pts && node_pos && idx, "null pointer");
    DFB_REQUIRE(k >= 1 && k <= DFB_MAX_K, "k=%d out of range [1,%d]", k, DFB_MAX_K);
    DFB_REQUIRE(n_nodes >= k, "n_nodes=%d < k=%d", n_nodes, k);
    DFB_REQUIRE(m >= 0, "negative point count");
    if (m == 0) return DFB_OK;
    const int64_t blocks = (m + 255) / 256;
    cudaStream_t s = (cudaStream_t)stream;
    if (k <= 4) knn_points_kernel<4><<<(unsigned)blocks, 256, 0, s>>>(pts, m, node_pos, n_nodes, k, idx);
    else knn_points_kernel<8><<<(unsigned)blocks, 256, 0, s>>>(pts, m, node_pos, n_nodes, k, idx);
    DFB_LAUNCH_CHECK("knn_points_kernel");
    return DFB_OK;
}
