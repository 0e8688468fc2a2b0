// tsdf.cu -- the three TSDF update passes (a1 volume-sampling, a2 rigid projective, a3 warped
// projective) as a fast HBM-streaming kernel plus a reference-exact clean-up kernel.
//
// Pass 1 (`*_fast_kernel`): one thread per voxel, z fastest (coalesced 128 B per warp for v, w and the
//   kNN table).  The fp32 tier classifies the voxel: certainly not updated (no memory traffic at all
//   beyond the kNN ids), certainly updated with the clamped value min(tdist, tl) = tdist (read v,w,
//   fp32 running average, write v,w), or uncertain (inside the truncation band, near the image border,
//   near a pixel-rounding boundary, ...) -> appended to a work list with one warp-aggregated atomic.
// Pass 2 (`*_exact_kernel`): a grid-stride kernel over the work list that evaluates the reference's
//   own arithmetic (dfb_math.h exact tier).  If the list overflowed it re-scans the volume instead.
#include <stdlib.h>

#include "common.h"
#include "dfb_params.h"
#include "dfb_brick.h"

using namespace dfb;

namespace {

// Volume data (v, w, kNN ids) is touched once per frame: streaming loads / stores (evict-first) keep the L1 / L2 lines for what
// IS re-used -- node records and depth pixels.  -DDFB_STREAM_HINTS=0 builds the plain variant (A/B in profiles/).
#ifndef DFB_STREAM_HINTS
#define DFB_STREAM_HINTS 1
#endif
__device__ __forceinline__ float4 ld_stream(const float* p) {
#if DFB_STREAM_HINTS
    return __ldcs(reinterpret_cast<const float4*>(p));
#else
    return *reinterpret_cast<const float4*>(p);
#endif
}
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void st_stream(float* p, float4 v) {
#if DFB_STREAM_HINTS
    __stcs(reinterpret_cast<float4*>(p), v);
#else
    *reinterpret_cast<float4*>(p) = v;
#endif
}
template <int KMAX>
__device__ __forceinline__ void load_ids(const uint16_t* knn, size_t i, int k, uint16_t* ids) {
    if (KMAX == 4) {
        if (k == 4) {
            const uint2 r = __ldg(reinterpret_cast<const uint2*>(knn) + i);
            ids[0] = r.x & 0xffff; ids[1] = r.x >> 16; ids[2] = r.y & 0xffff; ids[3] = r.y >> 16;
            return;
        }
    } else if (k == 8) {
        const uint4 r = __ldg(reinterpret_cast<const uint4*>(knn) + i);
        ids[0] = r.x & 0xffff; ids[1] = r.x >> 16; ids[2] = r.y & 0xffff; ids[3] = r.y >> 16;
        ids[4] = r.z & 0xffff; ids[5] = r.z >> 16; ids[6] = r.w & 0xffff; ids[7] = r.w >> 16;
        return;
    }
#pragma unroll
    for (int j = 0; j < KMAX; ++j)
        if (j < k) ids[j] = knn[i * (size_t)k + j];
}

// a deferred voxel that found the list full is marked in the overflow bitmap (the exact pass sweeps it after the list)
__device__ __forceinline__ void defer_voxel(uint32_t pos, uint32_t i, uint32_t* list, uint32_t capacity, uint32_t* overflow_bits) {
    if (pos < capacity) list[pos] = i;
    else if (overflow_bits) atomicOr(overflow_bits + (i >> 5), 1u << (i & 31));
}

__device__ __forceinline__ void push_uncertain(bool unc, uint32_t i, uint32_t* list, uint32_t capacity, uint32_t* counters,
                                               uint32_t* overflow_bits) {
    const unsigned ballot = __ballot_sync(0xffffffffu, unc);
    if (ballot == 0) return;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(ballot) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(counters, (uint32_t)__popc(ballot));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (unc) {
        const uint32_t pos = base + __popc(ballot & ((1u << lane) - 1u));
        defer_voxel(pos, i, list, capacity, overflow_bits);
    }
}

// ------------------------------------------------------------------------------------------------
// a2 / a3
// ------------------------------------------------------------------------------------------------
template <int KMAX>
__global__ void __launch_bounds__(256) proj_fast_kernel(const __grid_constant__ ProjParams P) {
    const int z = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int xs = blockIdx.z;
    const bool in = z < P.rz;
    const size_t i = ((size_t)xs * P.ry + y) * P.rz + (in ? z : 0);
    int cls = CLS_SKIP, m = 0, f = 0;
    if (in) {
        uint16_t ids[KMAX];
        if (!P.rigid) load_ids<KMAX>(P.knn, i, P.k, ids);
        cls = voxel_projective_classify<KMAX>(P, xs + P.x0, y, z, ids, &m, &f);
    }
    push_uncertain(in && cls == CLS_UNCERTAIN, (uint32_t)i, P.list, P.capacity, P.counters, P.overflow_bits);
    if (!in || cls == CLS_UNCERTAIN) return;
    if (m) {
        float v = P.tsdf[i], w = P.weight[i];
        const float sc = (float)P.scale;
        for (int vi = 0; vi < P.n_views; ++vi)
            if (m & (1 << vi)) clamp_update(v, w, P.tdist_f, P.wmax_f, sc);
        P.tsdf[i] = v;
        P.weight[i] = w;
    }
    if (P.mask_out) P.mask_out[i] = (uint8_t)m;
    if (P.frustum_out) P.frustum_out[i] = (uint8_t)f;
}


// ------------------------------------------------------------------------------------------------
// brick culling (dfb_brick.h): per-graph candidate sets, per-frame classification, streaming + mixed passes
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void brick_thread_coords(int b, int nby, int nbz, int& bxs, int& by, int& bz) {
    bz = b % nbz;
    const int t = b / nbz;
    by = t % nby;
    bxs = t / nby;
}

// work-list entries carry the brick coordinates packed 10/10/12 bits (x, y, z), so the per-voxel passes never divide
constexpr int BRICK_PACK_MAX_XY = 1024, BRICK_PACK_MAX_Z = 4096;
__device__ __forceinline__ uint32_t brick_pack(int bxs, int by, int bz) { return ((uint32_t)bxs << 22) | ((uint32_t)by << 12) | (uint32_t)bz; }
__device__ __forceinline__ void brick_unpack(uint32_t e, int& bxs, int& by, int& bz) {
    bxs = (int)(e >> 22);
    by = (int)((e >> 12) & 1023u);
    bz = (int)(e & 4095u);
}

// thread t of a 128-thread CTA owns 4 consecutive z voxels of one of the brick's 16 rows
__device__ __forceinline__ void brick_lane(int t, int& dx, int& dy, int& dz) {
    const int row = t >> 3;
    dx = row >> 2;
    dy = row & 3;
    dz = (t & 7) * 4;
}

// dirty8 (optional): flags of the 8^3 bricks of the kNN build whose rows changed (dfb_knn_update_volume); a brick / region none of
// whose voxels lies in a flagged 8^3 brick keeps its sets
__device__ __forceinline__ bool any_dirty8(const uint8_t* dirty8, int sx, int ry, int rz, int xlo, int xhi, int ylo, int yhi, int zlo, int zhi) {
    const int n8y = (ry + 7) / 8, n8z = (rz + 7) / 8;
    xhi = min(xhi, sx - 1); yhi = min(yhi, ry - 1); zhi = min(zhi, rz - 1);
    for (int a = xlo / 8; a <= xhi / 8; ++a)
        for (int b = ylo / 8; b <= yhi / 8; ++b)
            for (int c = zlo / 8; c <= zhi / 8; ++c)
                if (dirty8[((size_t)a * n8y + b) * n8z + c]) return true;
    return false;
}

__global__ void __launch_bounds__(128) brick_nodes_kernel(const uint16_t* knn, int k, int sx, int ry, int rz, int nby, int nbz,
                                                          uint16_t* brick_nodes, uint8_t* brick_count, uint32_t* brick_pairs, const uint8_t* dirty8) {
    __shared__ unsigned int set[64];
    __shared__ unsigned int list[BRICK_MAXC];
    __shared__ unsigned int pairs[BRICK_PAIR_WORDS];
    __shared__ int n_out, overflow;
    const int b = blockIdx.x;
    int bxs, by, bz;
    brick_thread_coords(b, nby, nbz, bxs, by, bz);
    if (dirty8 && !any_dirty8(dirty8, sx, ry, rz, bxs * BRICK_X, bxs * BRICK_X + BRICK_X - 1, by * BRICK_Y, by * BRICK_Y + BRICK_Y - 1, bz * BRICK_Z,
                              bz * BRICK_Z + BRICK_Z - 1))
        return;
    if (threadIdx.x < 64) set[threadIdx.x] = 0xffffffffu;
    if (threadIdx.x == 0) { n_out = 0; overflow = 0; }
    __syncthreads();
    int dx, dy, dz;
    brick_lane(threadIdx.x, dx, dy, dz);
    const int xs = bxs * BRICK_X + dx, y = by * BRICK_Y + dy;
    if (xs < sx && y < ry) {
        for (int q = 0; q < 4; ++q) {
            const int z = bz * BRICK_Z + dz + q;
            if (z >= rz) break;
            const size_t i = ((size_t)xs * ry + y) * rz + z;
            for (int j = 0; j < k; ++j) {
                const unsigned int id = knn[i * (size_t)k + j];
                unsigned int slot = (id * 2654435761u) >> 26;
                int probes = 0;
                while (true) {
                    const unsigned int old = atomicCAS(&set[slot], 0xffffffffu, id);
                    if (old == 0xffffffffu || old == id) break;
                    slot = (slot + 1) & 63;
                    if (++probes >= 64) { overflow = 1; break; }
                }
            }
        }
    }
    __syncthreads();
    if (threadIdx.x < BRICK_PAIR_WORDS) pairs[threadIdx.x] = 0u;
    if (threadIdx.x < 64 && set[threadIdx.x] != 0xffffffffu) {
        const int pos = atomicAdd(&n_out, 1);
        if (pos < BRICK_MAXC) {
            brick_nodes[(size_t)b * BRICK_MAXC + pos] = (uint16_t)set[threadIdx.x];
            list[pos] = set[threadIdx.x];
        }
    }
    __syncthreads();
    const bool ok = !overflow && n_out <= BRICK_MAXC;
    if (threadIdx.x == 0) brick_count[b] = ok ? (uint8_t)n_out : 255;
    if (!ok) return;
    // pairs of candidates that share a voxel: bit i*(i+1)/2 + j, j <= i (local indices into the brick's list)
    if (xs < sx && y < ry) {
        for (int q = 0; q < 4; ++q) {
            const int z = bz * BRICK_Z + dz + q;
            if (z >= rz) break;
            const size_t i = ((size_t)xs * ry + y) * rz + z;
            int loc[DFB_MAX_K];
            for (int j = 0; j < k; ++j) {
                const unsigned int id = knn[i * (size_t)k + j];
                int l = 0;
                while (l < n_out && list[l] != id) ++l;
                loc[j] = l;
            }
            for (int a = 0; a < k; ++a)
                for (int c2 = 0; c2 <= a; ++c2) {
                    const int hi = loc[a] > loc[c2] ? loc[a] : loc[c2], lo = loc[a] > loc[c2] ? loc[c2] : loc[a];
                    const int p = hi * (hi + 1) / 2 + lo;
                    atomicOr(&pairs[p >> 5], 1u << (p & 31));
                }
        }
    }
    __syncthreads();
    if (threadIdx.x < BRICK_PAIR_WORDS) brick_pairs[(size_t)b * BRICK_PAIR_WORDS + threadIdx.x] = pairs[threadIdx.x];
}


// ------------------------------------------------------------------------------------------------
// regions (dfb_brick.h): per-graph node/pair sets, per-frame reference map + deviation bound
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) region_build_kernel(const uint16_t* knn, int k, int sx, int ry, int rz, int nry, int nrz,
                                                           uint16_t* region_nodes, uint8_t* region_count, uint32_t* region_pairs, const uint8_t* dirty8) {
    __shared__ unsigned int hkey[256];
    __shared__ int hval[256];
    __shared__ unsigned int pairs[REGION_PAIR_WORDS];
    __shared__ int n_out, overflow;
    const int reg = blockIdx.x;
    const int rzi = reg % nrz, ryi = (reg / nrz) % nry, rxi = reg / (nrz * nry);
    if (dirty8 && !any_dirty8(dirty8, sx, ry, rz, rxi * REGION_X, rxi * REGION_X + REGION_X - 1, ryi * REGION_Y, ryi * REGION_Y + REGION_Y - 1,
                              rzi * REGION_Z, rzi * REGION_Z + REGION_Z - 1))
        return;
    for (int t = threadIdx.x; t < 256; t += blockDim.x) { hkey[t] = 0xffffffffu; hval[t] = -1; }
    for (int t = threadIdx.x; t < REGION_PAIR_WORDS; t += blockDim.x) pairs[t] = 0u;
    if (threadIdx.x == 0) { n_out = 0; overflow = 0; }
    __syncthreads();
    constexpr int NV = REGION_X * REGION_Y * REGION_Z;
    for (int v = threadIdx.x; v < NV; v += blockDim.x) {
        const int z = rzi * REGION_Z + (v % REGION_Z), y = ryi * REGION_Y + (v / REGION_Z) % REGION_Y, xs = rxi * REGION_X + v / (REGION_Z * REGION_Y);
        if (xs >= sx || y >= ry || z >= rz) continue;
        const size_t i = ((size_t)xs * ry + y) * rz + z;
        for (int j = 0; j < k; ++j) {
            const unsigned int id = knn[i * (size_t)k + j];
            unsigned int slot = (id * 2654435761u) >> 24;
            int probes = 0;
            while (true) {
                const unsigned int old = atomicCAS(&hkey[slot], 0xffffffffu, id);
                if (old == 0xffffffffu || old == id) break;
                slot = (slot + 1) & 255;
                if (++probes >= 256) { overflow = 1; break; }
            }
        }
    }
    __syncthreads();
    if (hkey[threadIdx.x] != 0xffffffffu) {
        const int pos = atomicAdd(&n_out, 1);
        hval[threadIdx.x] = pos;
        if (pos < REGION_MAXC) region_nodes[(size_t)reg * REGION_MAXC + pos] = (uint16_t)hkey[threadIdx.x];
    }
    __syncthreads();
    const bool ok = !overflow && n_out <= REGION_MAXC;
    if (threadIdx.x == 0) region_count[reg] = ok ? (uint8_t)n_out : 255;
    if (!ok) return;
    for (int v = threadIdx.x; v < NV; v += blockDim.x) {
        const int z = rzi * REGION_Z + (v % REGION_Z), y = ryi * REGION_Y + (v / REGION_Z) % REGION_Y, xs = rxi * REGION_X + v / (REGION_Z * REGION_Y);
        if (xs >= sx || y >= ry || z >= rz) continue;
        const size_t i = ((size_t)xs * ry + y) * rz + z;
        int loc[DFB_MAX_K];
        for (int j = 0; j < k; ++j) {
            const unsigned int id = knn[i * (size_t)k + j];
            unsigned int slot = (id * 2654435761u) >> 24;
            while (hkey[slot] != id) slot = (slot + 1) & 255;
            loc[j] = hval[slot];
        }
        for (int a = 0; a < k; ++a)
            for (int c2 = 0; c2 <= a; ++c2) {
                const int hi = loc[a] > loc[c2] ? loc[a] : loc[c2], lo = loc[a] > loc[c2] ? loc[c2] : loc[a];
                const int p = hi * (hi + 1) / 2 + lo;
                const unsigned int bit = 1u << (p & 31);
                if (!(pairs[p >> 5] & bit)) atomicOr(&pairs[p >> 5], bit);
            }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < REGION_PAIR_WORDS; t += blockDim.x) region_pairs[(size_t)reg * REGION_PAIR_WORDS + t] = pairs[t];
}

#ifndef DFB_REGION_THREADS
#define DFB_REGION_THREADS 96   // measured at 512^3: 64 -> 0.153, 96 -> 0.146, 128 -> 0.157 ms (region + brick classification)
#endif
static_assert(DFB_REGION_THREADS >= REGION_MAXC && DFB_REGION_THREADS % 32 == 0 && DFB_REGION_THREADS <= 128, "one thread per cached node");
#ifndef DFB_REGION_MINB
#define DFB_REGION_MINB 16   // CTAs per SM (40 registers): the kernel is a chain of dependent phases; measured at 512^3, region + brick
#endif                        // classification: unbounded (64 registers, 10 CTAs) 0.153 ms, 12 -> 0.150, 16 -> 0.141, 20 -> 0.144
__global__ void __launch_bounds__(DFB_REGION_THREADS, DFB_REGION_MINB) region_bounds_kernel(const __grid_constant__ ProjParams P, const uint16_t* region_nodes,
                                                            const uint8_t* region_count, const uint32_t* region_pairs, int nry, int nrz,
                                                            float* region_rec) {
    const float4* node_rec = P.node_rec;
    const int x0 = P.x0, sx = P.x1 - P.x0, ry = P.ry, rz = P.rz;
    __shared__ float q[REGION_MAXC][8];
    __shared__ float red[DFB_REGION_THREADS / 32][3];
    __shared__ int bad_s;
    __shared__ float As[REGION_MAXC][12];   // the nodes' own maps A(q_i) / |q_i|^2
    __shared__ float Pref_s[12];
    __shared__ int pre[REGION_PAIR_WORDS + 1];
    __shared__ uint16_t plist[REGION_MAXC * (REGION_MAXC + 1) / 2];
    // grid = (nrz, nry, nrx): no index division
    const int rzi = blockIdx.x, ryi = blockIdx.y, rxi = blockIdx.z;
    const int reg = (rxi * nry + ryi) * nrz + rzi;
    float* out = region_rec + (size_t)reg * REGION_REC_FLOATS;
    const int cnt = region_count[reg];
    if (cnt == 0 || cnt > REGION_MAXC) {
        if (threadIdx.x < REGION_REC_FLOATS) out[threadIdx.x] = 0.f;
        return;
    }
    const int xlo = rxi * REGION_X, ylo = ryi * REGION_Y, zlo = rzi * REGION_Z;
    const int xhi = min(xlo + REGION_X, sx) - 1, yhi = min(ylo + REGION_Y, ry) - 1, zhi = min(zlo + REGION_Z, rz) - 1;
    const float c[3] = {0.5f * (xlo + xhi) + (float)x0, 0.5f * (ylo + yhi), 0.5f * (zlo + zhi)};
    const float h[3] = {0.5f * (xhi - xlo), 0.5f * (yhi - ylo), 0.5f * (zhi - zlo)};
    if (threadIdx.x == 0) bad_s = 0;
    __syncthreads();
    bool bad = false;
    const uint16_t* ids = region_nodes + (size_t)reg * REGION_MAXC;
    const int npairs = cnt * (cnt + 1) / 2, nwords = (npairs + 31) >> 5;
    const uint32_t* pm = region_pairs + (size_t)reg * REGION_PAIR_WORDS;
    if (threadIdx.x < cnt) {
        const int id = ids[threadIdx.x];
        const float4 r0 = node_rec[3 * (size_t)id], r1 = node_rec[3 * (size_t)id + 1], r2 = node_rec[3 * (size_t)id + 2];
        q[threadIdx.x][0] = r1.x; q[threadIdx.x][1] = r1.y; q[threadIdx.x][2] = r1.z; q[threadIdx.x][3] = r1.w;
        q[threadIdx.x][4] = r2.x; q[threadIdx.x][5] = r2.y; q[threadIdx.x][6] = r2.z; q[threadIdx.x][7] = r2.w;
        // every blend weight of every voxel must be a normal float32 in the reference (see dfb_voxel.h blend_warp_fast)
        const float ddx = fabsf(c[0] - r0.x) + h[0], ddy = fabsf(c[1] - r0.y) + h[1], ddz = fabsf(c[2] - r0.z) + h[2];
        if (!((ddx * ddx + ddy * ddy + ddz * ddz) * r0.w > -125.f)) bad = true;
        if (!node_affine_normalised(q[threadIdx.x], As[threadIdx.x])) bad = true;
    }
    if (threadIdx.x >= DFB_REGION_THREADS - 32) {
        // last warp: exclusive prefix of the pair mask's popcounts, so that the pairs that do co-occur can be dealt out densely
        const int lane = threadIdx.x & 31;
        int carry = 0;
        for (int w0 = 0; w0 < nwords; w0 += 32) {
            const int w = w0 + lane;
            uint32_t m = w < nwords ? pm[w] : 0u;
            if (w == nwords - 1 && (npairs & 31)) m &= (1u << (npairs & 31)) - 1u;
            const int cntw = __popc(m);
            int inc = cntw;
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += v;
            }
            if (w < nwords) pre[w] = carry + inc - cntw;
            carry += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) pre[nwords] = carry;
    }
    __syncthreads();
    for (int w = threadIdx.x; w < nwords; w += blockDim.x) {
        uint32_t m = pm[w];
        if (w == nwords - 1 && (npairs & 31)) m &= (1u << (npairs & 31)) - 1u;
        int o = pre[w];
        while (m) {
            const int bit = __ffs(m) - 1;
            m &= m - 1;
            plist[o++] = (uint16_t)(32 * w + bit);
        }
    }
    __syncthreads();
    // reference map = mean of the nodes' own maps (dfb_brick.h region_reference_map), summed in list order
    if (threadIdx.x < 12) {
        float a = 0.f;
        for (int t = 0; t < cnt; ++t) a += As[t][threadIdx.x];
        Pref_s[threadIdx.x] = a * (1.0f / (float)cnt);
    }
    __syncthreads();
    float Pref[12];
    for (int t = 0; t < 12; ++t) Pref[t] = Pref_s[t];
    float dev[3] = {0.f, 0.f, 0.f};
    const int nset = pre[nwords];
    for (int t = threadIdx.x; t < nset; t += blockDim.x) {
        const int p = plist[t];
        int i = (int)((sqrtf(8.f * (float)p + 1.f) - 1.f) * 0.5f);
        while (i * (i + 1) / 2 > p) --i;
        while ((i + 1) * (i + 2) / 2 <= p) ++i;
        const int j = p - i * (i + 1) / 2;
        float qi[8], qj[8];
        for (int t2 = 0; t2 < 8; ++t2) { qi[t2] = q[i][t2]; qj[t2] = q[j][t2]; }
        if (!region_pair_bound(qi, qj, i == j, Pref, c, h, dev)) bad = true;
    }
    if (bad) atomicOr(&bad_s, 1);
    for (int r = 0; r < 3; ++r)
        for (int o = 16; o > 0; o >>= 1) dev[r] = fmaxf(dev[r], __shfl_xor_sync(0xffffffffu, dev[r], o));
    if ((threadIdx.x & 31) == 0)
        for (int r = 0; r < 3; ++r) red[threadIdx.x >> 5][r] = dev[r];
    __syncthreads();
    if (threadIdx.x < 32) {
        // warp 0: the record, and -- with (P_ref, D_R) in hand -- one attempt to classify the region as a whole; every
        // brick of a region that is SKIP / CLAMP in its entirety inherits that class without any work of its own
        float rr[16];
        for (int t = 0; t < 12; ++t) rr[t] = Pref[t];
        for (int r = 0; r < 3; ++r) {
            float mx = red[0][r];
            for (int wv = 1; wv < DFB_REGION_THREADS / 32; ++wv) mx = fmaxf(mx, red[wv][r]);
            rr[12 + r] = mx;
        }
        const bool valid = bad_s == 0;
        BrickClass rc = brick_class_all_mixed(P.n_views);
        if (valid) {
            Box3 bx;
            region_box(rr, c, h, P.coord_mag, bx);
            rc = box_classify_views(P, bx, REGION_MAX_RECT, WarpCtx());
        }
        rr[15] = region_code(valid, rc);
        if (threadIdx.x < 16) {
            const int l = threadIdx.x;
            out[l] = l < 12 ? Pref_s[l] : l == 12 ? rr[12] : l == 13 ? rr[13] : l == 14 ? rr[14] : rr[15];
        }
    } else if (threadIdx.x < 64) {
        // warp 1, meanwhile: the update pass's quad pre-test works with lpos = M [x, 1] +- d of view 0 (one-view frames; dfb_brick.h
        // quad_view_setup) -- composed once per region here instead of once per brick layer there, one entry per lane
        const int l = threadIdx.x - 32;
        const float* T = P.vf[0].T;
        float val = 0.f;
        if (l < 12) {
            const int r = l >> 2, cc = l & 3;
            val = T[4 * r] * Pref_s[cc] + T[4 * r + 1] * Pref_s[4 + cc] + T[4 * r + 2] * Pref_s[8 + cc] + (cc == 3 ? T[4 * r + 3] : 0.f);
        } else if (l < 15) {
            const int r = l - 12;
            float D[3];
            for (int a = 0; a < 3; ++a) {
                float mx = red[0][a];
                for (int wv = 1; wv < DFB_REGION_THREADS / 32; ++wv) mx = fmaxf(mx, red[wv][a]);
                D[a] = mx + 2e-3f + 2e-6f * P.coord_mag;
            }
            val = fabsf(T[4 * r]) * D[0] + fabsf(T[4 * r + 1]) * D[1] + fabsf(T[4 * r + 2]) * D[2] + 8e-6f * P.coord_mag + 1e-3f;
        }
        if (l < 16) out[16 + l] = val;
    }
}

// 8 lanes per brick, four bricks per warp (lanes split the candidate-node pairs and the depth pixels; dfb_brick.h)
#ifdef DFB_CLASSIFY_MINB   // measured: 12 / 16 CTAs per SM (40 / 32 registers) do not beat the compiler's own 56 registers
#define DFB_CLASSIFY_BOUNDS __launch_bounds__(128, DFB_CLASSIFY_MINB)
#else
#define DFB_CLASSIFY_BOUNDS __launch_bounds__(128)
#endif
template <int CLASSIFY_G>
__global__ void DFB_CLASSIFY_BOUNDS brick_classify_kernel(const __grid_constant__ ProjParams P, const uint16_t* brick_nodes,
                                                             const uint8_t* brick_count, const uint32_t* brick_pairs, const float* region_rec,
                                                             int nbx, int nby, int nbz, uint8_t* cls_out, uint32_t* stream_list,
                                                             uint32_t* mixed_list) {
    const int nb = nbx * nby * nbz;
    const int gl = threadIdx.x & (CLASSIFY_G - 1);
    const int ngroups = (gridDim.x * blockDim.x) / CLASSIFY_G;
    // bricks are visited region by region (a region = 4 x 4 x 1 bricks): the groups of a warp then share their region, and a
    // warp whose region was settled as a whole leaves together instead of idling next to groups that still have work
    constexpr int RBX = REGION_X / BRICK_X, RBY = REGION_Y / BRICK_Y;
    static_assert(REGION_Z == BRICK_Z && RBX * RBY == 16, "region = 4 x 4 x 1 bricks");
    const int nrx = (nbx + RBX - 1) / RBX, nry = (nby + RBY - 1) / RBY;
    const int total = nrx * nry * nbz * 16;
    for (int g = (blockIdx.x * blockDim.x + threadIdx.x) / CLASSIFY_G; g < total; g += ngroups) {
        const int reg = g >> 4, local = g & 15;
        const int bz = reg % nbz, t = reg / nbz;
        const int bxs = (t / nry) * RBX + (local >> 2), by = (t % nry) * RBY + (local & 3);
        if (bxs >= nbx || by >= nby) continue;
        const int b = (bxs * nby + by) * nbz + bz;
        const BrickClass bc = brick_classify(P, brick_nodes, brick_count, brick_pairs, region_rec, nby, nbz, bxs, by, bz, GroupCtx<CLASSIFY_G>());
        if (gl == 0) {
            // four planes of nb bytes: class (0xFF = MIXED), frustum bits, open views, CLAMP bits of the settled views
            cls_out[b] = (uint8_t)bc.cls;
            cls_out[nb + b] = (uint8_t)bc.frus;
            cls_out[2 * nb + b] = (uint8_t)bc.mixed;
            cls_out[3 * nb + b] = (uint8_t)bc.clamp;
            // (one list reservation per warp instead of per brick was measured: no difference, 0.1461 vs 0.1466 ms)
            if (bc.mixed) mixed_list[atomicAdd(P.counters + 3, 1u)] = brick_pack(bxs, by, bz);
            else if (bc.clamp != 0 || (P.frustum_out != nullptr && bc.frus != 0)) stream_list[atomicAdd(P.counters + 2, 1u)] = brick_pack(bxs, by, bz);
        }
    }
}

// CLAMP bricks: v' = (scale*v*w + tdist)/(scale*(w+1)), w' = min(w+1, wmax) once per view bit -- no warp, no kNN read.
template <bool ONEVIEW = false>
__device__ __forceinline__ void stream_brick(const ProjParams& P, int nb, int nby, int nbz, const uint8_t* cls, uint32_t entry, int dx, int dy,
                                             int dz, float sc, bool vec) {
    int bxs, by, bz;
    brick_unpack(entry, bxs, by, bz);
    const int b = (bxs * nby + by) * nbz + bz;
    const int m = cls[b];
    const int xs = bxs * BRICK_X + dx, y = by * BRICK_Y + dy, z = bz * BRICK_Z + dz;
    if (xs >= P.x1 - P.x0 || y >= P.ry || z >= P.rz) return;
    const uint32_t i = ((uint32_t)xs * (uint32_t)P.ry + (uint32_t)y) * (uint32_t)P.rz + (uint32_t)z;   // a slab holds < 2^32 voxels
    const bool want_masks = P.mask_out != nullptr || P.frustum_out != nullptr;
    if (vec) {
        if (m) {
            float4 v = ld_stream(P.tsdf + i);
            float4 w = ld_stream(P.weight + i);
            if (ONEVIEW) {
                clamp_update(v.x, w.x, P.tdist_f, P.wmax_f, sc);
                clamp_update(v.y, w.y, P.tdist_f, P.wmax_f, sc);
                clamp_update(v.z, w.z, P.tdist_f, P.wmax_f, sc);
                clamp_update(v.w, w.w, P.tdist_f, P.wmax_f, sc);
            } else {
                for (int vi = 0; vi < P.n_views; ++vi)
                    if (m & (1 << vi)) {
                        clamp_update(v.x, w.x, P.tdist_f, P.wmax_f, sc);
                        clamp_update(v.y, w.y, P.tdist_f, P.wmax_f, sc);
                        clamp_update(v.z, w.z, P.tdist_f, P.wmax_f, sc);
                        clamp_update(v.w, w.w, P.tdist_f, P.wmax_f, sc);
                    }
            }
            st_stream(P.tsdf + i, v);
            st_stream(P.weight + i, w);
        }
        if (want_masks) {
            const int fr = cls[nb + b];
            if (P.mask_out) *reinterpret_cast<uchar4*>(P.mask_out + i) = make_uchar4(m, m, m, m);
            if (P.frustum_out) *reinterpret_cast<uchar4*>(P.frustum_out + i) = make_uchar4(fr, fr, fr, fr);
        }
    } else {
        const int fr = cls[nb + b];
        for (int q = 0; q < 4 && z + q < P.rz; ++q) {
            if (m) {
                float v = P.tsdf[i + q], w = P.weight[i + q];
                for (int vi = 0; vi < P.n_views; ++vi)
                    if (m & (1 << vi)) clamp_update(v, w, P.tdist_f, P.wmax_f, sc);
                P.tsdf[i + q] = v;
                P.weight[i + q] = w;
            }
            if (P.mask_out) P.mask_out[i + q] = (uint8_t)m;
            if (P.frustum_out) P.frustum_out[i + q] = (uint8_t)fr;
        }
    }
}

// MIXED bricks: the per-voxel fast tier (classify, clamped update, defer the rest to the exact pass).
// Edge bricks (cut by the volume boundary, or rz not a multiple of 4): one guarded voxel at a time.
template <int KMAX, bool EXACTK, bool ONEVIEW, class Rec>
__device__ __forceinline__ void mixed_brick_edge(const ProjParams& P, int bxs, int by, int bz, int dx, int dy, int dz, float sc, int views, int m0,
                                                 int f0, const Rec rec) {
    const bool want_masks = P.mask_out != nullptr || P.frustum_out != nullptr;
    const int xs = bxs * BRICK_X + dx, y = by * BRICK_Y + dy, z0 = bz * BRICK_Z + dz;
    const bool row_in = xs < P.x1 - P.x0 && y < P.ry;
    for (int q = 0; q < 4; ++q) {
        const int z = z0 + q;
        const bool in = row_in && z < P.rz;
        const size_t i = in ? ((size_t)xs * P.ry + y) * P.rz + z : 0;
        int cls = CLS_SKIP, m = 0, f = 0;
        if (in) {
            uint16_t ids[KMAX];
            if (!P.rigid) load_ids<KMAX>(P.knn, i, EXACTK ? KMAX : P.k, ids);
            cls = voxel_projective_classify_rec<KMAX, EXACTK, ONEVIEW>(P, xs + P.x0, y, z, ids, &m, &f, views, m0, f0, rec);
        }
        push_uncertain(in && cls == CLS_UNCERTAIN, (uint32_t)i, P.list, P.capacity, P.counters, P.overflow_bits);
        if (!in || cls == CLS_UNCERTAIN) continue;
        if (m) {
            float v = P.tsdf[i], w = P.weight[i];
            for (int vi = 0; vi < P.n_views; ++vi)
                if (m & (1 << vi)) clamp_update(v, w, P.tdist_f, P.wmax_f, sc);
            P.tsdf[i] = v;
            P.weight[i] = w;
        }
        if (want_masks) {
            if (P.mask_out) P.mask_out[i] = (uint8_t)m;
            if (P.frustum_out) P.frustum_out[i] = (uint8_t)f;
        }
    }
}

// ---- MIXED bricks of the production pass: quad pre-test + a per-warp queue of open voxels ----------------------------------------
// A warp owns one x-layer of a brick (4 rows x 32 z = 128 voxels, one quad of four z-consecutive voxels per lane).
//  1. Every lane runs the quad pre-test (dfb_brick.h "quads": the region's reference map + deviation bound against the one to four
//     depth pixels the quad can project to): each voxel is settled (SKIP / CLAMP), certainly inside the band (-> the exact pass'
//     work list, without a look at the nodes), or open.  Settled voxels are folded into (v, w) right away -- float4 traffic, and
//     only for quads that update anything.
//  2. Open voxels go into the warp's queue in shared memory (ballot compaction).  Whenever the queue holds 32 entries, the warp
//     takes them through the pointwise DQB tier with every lane busy; what is left over stays queued for the next brick, so the
//     expensive tier always runs at full width however few voxels a brick leaves open.  No block-level barrier anywhere.
// Measured at 512^3 (profiles/): 74 % of the voxels of MIXED bricks never reach the DQB tier.
constexpr int WQ_CAP = 96;                          // < 32 left over + two appends of <= 32 (the quad is queued in two halves)
constexpr size_t WQ_BYTES = 2 * WQ_CAP * sizeof(uint32_t);
struct WarpQueue {
    uint32_t* vox;   // slab-linear voxel index
    uint32_t* xy;    // slab x << 16 | y
    int n;           // warp-uniform
};

// 32 (or, draining, fewer) queued voxels through the pointwise tier: classify, fold the clamped update into (v, w), defer the rest
template <int KMAX, bool EXACTK, bool ONEVIEW, class Rec>
__device__ __forceinline__ void queue_process(const ProjParams& P, const uint8_t* cls, int nb, int nby, int nbz, float sc, const Rec rec,
                                              WarpQueue& Q, int count) {
    const int lane = threadIdx.x & 31;
    const bool act = lane < count;
    const int e = Q.n - count + lane;
    uint32_t i = 0;
    int c = CLS_SKIP, mv = 0, fv = 0;
    float v = 0.f, w = 0.f;
    if (act) {
        i = Q.vox[e];
        // requested before the classification (some 300 instructions) so that the update below does not wait for them
        v = P.tsdf[i];
        w = P.weight[i];
        const uint32_t xy = Q.xy[e];
        const int xs = (int)(xy >> 16), y = (int)(xy & 0xffffu);
        const int z = (int)(i - ((uint32_t)xs * (uint32_t)P.ry + (uint32_t)y) * (uint32_t)P.rz);
        int views = 0xff, m0 = 0, f0 = 0;
        if (!ONEVIEW) {   // what the brick's box test settled (several views)
            const int b = ((xs / BRICK_X) * nby + y / BRICK_Y) * nbz + z / BRICK_Z;
            f0 = cls[nb + b]; views = cls[2 * nb + b]; m0 = cls[3 * nb + b];
        }
        uint16_t ids[KMAX];
        load_ids<KMAX>(P.knn, i, EXACTK ? KMAX : P.k, ids);
        c = voxel_projective_classify_rec<KMAX, EXACTK, ONEVIEW>(P, xs + P.x0, y, z, ids, &mv, &fv, views, m0, f0, rec);
    }
    push_uncertain(act && c == CLS_UNCERTAIN, i, P.list, P.capacity, P.counters, P.overflow_bits);
    if (act) {
        if (c == CLS_UNCERTAIN) { mv = 0; fv = 0; }
        if (mv) {
            if (ONEVIEW) clamp_update(v, w, P.tdist_f, P.wmax_f, sc);
            else
                for (int vi = 0; vi < P.n_views; ++vi)
                    if (mv & (1 << vi)) clamp_update(v, w, P.tdist_f, P.wmax_f, sc);
            P.tsdf[i] = v;
            P.weight[i] = w;
        }
        if (P.mask_out) P.mask_out[i] = (uint8_t)mv;
        if (P.frustum_out) P.frustum_out[i] = (uint8_t)fv;
    }
    Q.n -= count;
    __syncwarp();
}

// One x-layer (dx = layer) of an interior MIXED brick whose region carries a deviation bound `rr`: pre-test, settled voxels folded
// into (v, w), band voxels deferred.  Returns the lane's open voxels (bit q = voxel z0 + q goes through the pointwise tier).
template <bool ONEVIEW>
__device__ __forceinline__ uint32_t mixed_layer_quads(const ProjParams& P, const float* rr, int xs, int y, int z0, uint32_t i0, float sc, int views, int m0,
                                                      int f0) {
    const int lane = threadIdx.x & 31;
    // lpos = M [x, 1] +- d per view.  One view: region_bounds_kernel left view 0's (M, d) in the region record (one 128-bit
    // broadcast load per four entries); several views: lane l < 15 composes entry l of view v, the warp shares them by shuffle
    QuadView qv[ONEVIEW ? 1 : DFB_MAX_VIEWS];
    const int nv = ONEVIEW ? 1 : P.n_views;
    if (ONEVIEW) {
        const float4* q4 = reinterpret_cast<const float4*>(rr + 16);
        const float4 a = __ldg(q4), b = __ldg(q4 + 1), c = __ldg(q4 + 2), d = __ldg(q4 + 3);
        qv[0].M[0] = a.x; qv[0].M[1] = a.y; qv[0].M[2] = a.z; qv[0].M[3] = a.w;
        qv[0].M[4] = b.x; qv[0].M[5] = b.y; qv[0].M[6] = b.z; qv[0].M[7] = b.w;
        qv[0].M[8] = c.x; qv[0].M[9] = c.y; qv[0].M[10] = c.z; qv[0].M[11] = c.w;
        qv[0].d[0] = d.x; qv[0].d[1] = d.y; qv[0].d[2] = d.z;
    }
    for (int v = 0; v < (ONEVIEW ? 0 : nv); ++v) {
        if (!((views >> v) & 1)) continue;
        float val = 0.f;
        {
            const float* T = P.vf[v].T;
            if (lane < 12) {
                const int r = lane >> 2, c = lane & 3;
                val = T[4 * r] * rr[c] + T[4 * r + 1] * rr[4 + c] + T[4 * r + 2] * rr[8 + c] + (c == 3 ? T[4 * r + 3] : 0.f);
            } else if (lane < 15) {
                const int r = lane - 12;
                const float pad = 2e-3f + 2e-6f * P.coord_mag;   // the inflation region_box applies (quad_view_setup)
                val = fabsf(T[4 * r]) * (rr[12] + pad) + fabsf(T[4 * r + 1]) * (rr[13] + pad) + fabsf(T[4 * r + 2]) * (rr[14] + pad) +
                      8e-6f * P.coord_mag + 1e-3f;
            }
        }
#pragma unroll
        for (int t = 0; t < 12; ++t) qv[v].M[t] = __shfl_sync(0xffffffffu, val, t);
#pragma unroll
        for (int t = 0; t < 3; ++t) qv[v].d[t] = __shfl_sync(0xffffffffu, val, 12 + t);
    }
    int st[4], mq[4], fq[4];
    quad_pretest(P, qv, xs + P.x0, y, z0, 4, ONEVIEW ? 1 : views, m0, f0, st, mq, fq);
    uint32_t om = 0, bm = 0;
    bool any_upd = false;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        if (st[q] == QV_BAND) { mq[q] = 0; fq[q] = 0; bm |= 1u << q; }
        if (st[q] == QV_OPEN) om |= 1u << q;
        any_upd |= st[q] == QV_SKIP && mq[q] != 0;
    }
    // settled voxels
    if (any_upd) {
        const float4 v4 = ld_stream(P.tsdf + i0), w4 = ld_stream(P.weight + i0);
        float v[4] = {v4.x, v4.y, v4.z, v4.w}, w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (st[q] != QV_SKIP || !mq[q]) continue;
            if (ONEVIEW) clamp_update(v[q], w[q], P.tdist_f, P.wmax_f, sc);
            else
                for (int vi = 0; vi < P.n_views; ++vi)
                    if (mq[q] & (1 << vi)) clamp_update(v[q], w[q], P.tdist_f, P.wmax_f, sc);
        }
        if (!om) {   // band voxels keep their value (the exact pass runs after this kernel): the whole quad can be stored
            st_stream(P.tsdf + i0, make_float4(v[0], v[1], v[2], v[3]));
            st_stream(P.weight + i0, make_float4(w[0], w[1], w[2], w[3]));
        } else {     // open voxels of this quad are written later by whichever lane takes them from the queue
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (st[q] == QV_SKIP && mq[q]) { P.tsdf[i0 + q] = v[q]; P.weight[i0 + q] = w[q]; }
        }
    }
    if (P.mask_out || P.frustum_out) {
        if (!om) {
            if (P.mask_out) *reinterpret_cast<uchar4*>(P.mask_out + i0) = make_uchar4(mq[0], mq[1], mq[2], mq[3]);
            if (P.frustum_out) *reinterpret_cast<uchar4*>(P.frustum_out + i0) = make_uchar4(fq[0], fq[1], fq[2], fq[3]);
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (st[q] != QV_OPEN) {
                    if (P.mask_out) P.mask_out[i0 + q] = (uint8_t)mq[q];
                    if (P.frustum_out) P.frustum_out[i0 + q] = (uint8_t)fq[q];
                }
        }
    }
    // certainly inside the band: deferred without a look at the nodes
    if (__any_sync(0xffffffffu, bm != 0)) {
        const int mine = __popc(bm);
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += up;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(P.counters, (uint32_t)total);
        base = __shfl_sync(0xffffffffu, base, 0) + (uint32_t)(incl - mine);
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if ((bm >> q) & 1u) defer_voxel(base++, i0 + q, P.list, P.capacity, P.overflow_bits);
    }
    return om;
}

// region record of brick (bxs, by, bz) when it carries a valid deviation bound, else nullptr
__device__ __forceinline__ const float* brick_region_rec(const ProjParams& P, const float* region_rec, int bxs, int by, int bz) {
    if (!region_rec || P.rigid) return nullptr;
    const int nry = (P.ry + REGION_Y - 1) / REGION_Y, nrz = (P.rz + REGION_Z - 1) / REGION_Z;
    const uint32_t reg = ((uint32_t)(bxs * BRICK_X / REGION_X) * (uint32_t)nry + (uint32_t)(by * BRICK_Y / REGION_Y)) * (uint32_t)nrz + (uint32_t)(bz * BRICK_Z / REGION_Z);
    const float* rr = region_rec + (size_t)reg * REGION_REC_FLOATS;
    return rr[15] > 0.5f ? rr : nullptr;
}

// One warp's x-layer of MIXED brick `entry`.  Edge bricks (cut by the volume boundary) are finished here, one guarded voxel at a
// time; interior bricks return the lane's open voxels (all four when the region carries no deviation bound).
template <int KMAX, bool EXACTK, bool ONEVIEW, class Rec>
__device__ __forceinline__ uint32_t mixed_layer(const ProjParams& P, const float* region_rec, int nb, int nby, int nbz, const uint8_t* cls, uint32_t entry,
                                                int t128, float sc, const Rec rec, uint32_t* i0_out, uint32_t* xy_out) {
    int bxs, by, bz;
    brick_unpack(entry, bxs, by, bz);
    int views = 0xff, m0 = 0, f0 = 0;
    if (!ONEVIEW) {
        const int b = (bxs * nby + by) * nbz + bz;
        f0 = cls[nb + b]; views = cls[2 * nb + b]; m0 = cls[3 * nb + b];
    }
    int dx, dy, dz;
    brick_lane(t128, dx, dy, dz);
    const bool full = (P.rz & 3) == 0 && (bxs + 1) * BRICK_X <= P.x1 - P.x0 && (by + 1) * BRICK_Y <= P.ry && (bz + 1) * BRICK_Z <= P.rz;
    if (!full) {
        mixed_brick_edge<KMAX, EXACTK, ONEVIEW>(P, bxs, by, bz, dx, dy, dz, sc, views, m0, f0, rec);
        return 0u;
    }
    const int xs = bxs * BRICK_X + dx, y = by * BRICK_Y + dy, z0 = bz * BRICK_Z + dz;
    const uint32_t i0 = (uint32_t)(((size_t)xs * P.ry + y) * P.rz + z0);
    *i0_out = i0;
    *xy_out = ((uint32_t)xs << 16) | (uint32_t)y;
    // several views: the brick's box test has already settled most of them and the per-view quad test does not pay (measured, 8 views
    // k = 8 at 256^3: 0.61 ms with the queue alone, 0.81 ms with the pre-test) -- every voxel goes to the pointwise tier
    const float* rr = ONEVIEW ? brick_region_rec(P, region_rec, bxs, by, bz) : nullptr;
    if (!rr) return 0xfu;
    return mixed_layer_quads<ONEVIEW>(P, rr, xs, y, z0, i0, sc, views, m0, f0);
}

// The loop of the production pass for one 128-thread group (t128 = thread within the group): list entry t = one MIXED brick (its
// four x-layers go to the group's four warps) + the group's share of CLAMP bricks.  The pointwise tier has ONE call site: after
// every brick the warp appends its open voxels (two voxels of every quad at a time, the queue holds 96) and runs full rounds of 32;
// one extra pass after the last brick drains the queue.  (Inlining the tier at several sites made the kernel instruction-cache bound:
// stall_no_instruction 27 % of all samples, profiles/.)
template <int KMAX, bool EXACTK, bool ONEVIEW, class Rec>
__device__ __forceinline__ void update_body(const ProjParams& P, const float* region_rec, int nbx, int nby, int nbz, const uint8_t* cls,
                                            const uint32_t* stream_list, uint32_t cnt_s, const uint32_t* mixed_list, uint32_t cnt_m, uint32_t first,
                                            uint32_t stride, uint32_t n_t, int t128, const Rec rec, WarpQueue& Q) {
    const int nb = nbx * nby * nbz;
    const uint32_t share = n_t ? (cnt_s + n_t - 1) / n_t : 0u;
    const int lane = threadIdx.x & 31;
    int dx, dy, dz;
    brick_lane(t128, dx, dy, dz);
    const float sc = (float)P.scale;
    const bool vec = (P.rz & 3) == 0;
    uint32_t n_dqb = 0;
    for (uint32_t t = first;; t += stride) {
        const bool last = !(t < n_t);          // uniform across the group
        uint32_t om = 0, i0 = 0, xy = 0;
        if (!last) {
            if (t < cnt_m) om = mixed_layer<KMAX, EXACTK, ONEVIEW>(P, region_rec, nb, nby, nbz, cls, mixed_list[t], t128, sc, rec, &i0, &xy);
            const uint32_t s1 = (t + 1) * share < cnt_s ? (t + 1) * share : cnt_s;
            for (uint32_t s = t * share; s < s1; ++s) stream_brick<ONEVIEW>(P, nb, nby, nbz, cls, stream_list[s], dx, dy, dz, sc, vec);
        }
        const bool any = __any_sync(0xffffffffu, om != 0);
        if (om) {   // the quad's node ids, values and weights are needed a queue round from now, by some other lane of this warp
            prefetch_l1(P.knn + (size_t)i0 * P.k);
            prefetch_l1(P.tsdf + i0);
            prefetch_l1(P.weight + i0);
        }
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            if (any) {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int q = 2 * half + j;
                    const bool open = (om >> q) & 1u;
                    const unsigned bo = __ballot_sync(0xffffffffu, open);
                    if (open) {
                        const int pos = Q.n + __popc(bo & ((1u << lane) - 1u));
                        Q.vox[pos] = i0 + (uint32_t)q;
                        Q.xy[pos] = xy;
                    }
                    Q.n += __popc(bo);
                    n_dqb += __popc(bo);
                }
                __syncwarp();
            }
            const int thr = last ? 1 : 32;
            while (Q.n >= thr) queue_process<KMAX, EXACTK, ONEVIEW>(P, cls, nb, nby, nbz, sc, rec, Q, Q.n < 32 ? Q.n : 32);
        }
        if (last) break;
    }
    if (lane == 0 && n_dqb) atomicAdd(P.counters + 4, n_dqb);   // statistics: voxels that went through the pointwise DQB tier
}

__global__ void __launch_bounds__(128) brick_stream_kernel(const __grid_constant__ ProjParams P, int nbx, int nby, int nbz,
                                                           const uint8_t* cls, const uint32_t* list) {
    const uint32_t count = P.counters[2];
    int dx, dy, dz;
    brick_lane(threadIdx.x, dx, dy, dz);
    for (uint32_t t = blockIdx.x; t < count; t += gridDim.x)
        stream_brick(P, nbx * nby * nbz, nby, nbz, cls, list[t], dx, dy, dz, (float)P.scale, (P.rz & 3) == 0);
}

template <int KMAX, bool EXACTK, bool ONEVIEW>
__global__ void __launch_bounds__(128) brick_mixed_kernel(const __grid_constant__ ProjParams P, int nbx, int nby, int nbz,
                                                          const uint8_t* cls, const uint32_t* list, const float* region_rec) {
    __shared__ uint32_t s_queue[4][2 * WQ_CAP];
    WarpQueue Q = {s_queue[threadIdx.x >> 5], s_queue[threadIdx.x >> 5] + WQ_CAP, 0};
    const uint32_t count = P.counters[3];
    update_body<KMAX, EXACTK, ONEVIEW>(P, region_rec, nbx, nby, nbz, cls, nullptr, 0u, list, count, blockIdx.x, gridDim.x, count, threadIdx.x,
                                       RecGlobal{P.node_rec}, Q);
}

// Production pass: MIXED and CLAMP bricks in ONE persistent launch.  Every CTA alternates between one MIXED brick
// (issue-bound per-voxel tier) and its share of CLAMP bricks (HBM-bound streaming), so both kinds of work are resident
// on every SM at the same time and the streaming traffic hides under the arithmetic.
// 8 CTAs of 128 threads per SM (64 registers): measured best against 5, 6 and 10 (profiles/r1_ncu_full_summary.md)
#ifndef DFB_UPDATE_MINB
#define DFB_UPDATE_MINB 8
#endif
#define DFB_UPDATE_BOUNDS __launch_bounds__(128, DFB_UPDATE_MINB)
template <int KMAX, bool EXACTK, bool ONEVIEW>
__global__ void DFB_UPDATE_BOUNDS brick_update_kernel(const __grid_constant__ ProjParams P, int nbx, int nby, int nbz,
                                                           const uint8_t* cls, const uint32_t* stream_list, const uint32_t* mixed_list,
                                                           const float* region_rec) {
    __shared__ uint32_t s_queue[4][2 * WQ_CAP];
    WarpQueue Q = {s_queue[threadIdx.x >> 5], s_queue[threadIdx.x >> 5] + WQ_CAP, 0};
    const uint32_t cnt_s = P.counters[2], cnt_m = P.counters[3];
    const uint32_t n_t = cnt_m > gridDim.x ? cnt_m : gridDim.x;
    update_body<KMAX, EXACTK, ONEVIEW>(P, region_rec, nbx, nby, nbz, cls, stream_list, cnt_s, mixed_list, cnt_m, blockIdx.x, gridDim.x, n_t, threadIdx.x,
                                       RecGlobal{P.node_rec}, Q);
}

// The same pass with the WHOLE node table resident in shared memory (north_star: "node parameters ... staged in shared memory or
// via TMA"): one persistent CTA of 1024 threads per SM copies the packed records (48 B per node; 191 KB at 3974 nodes) with
// cp.async.bulk (the TMA unit, one mbarrier) and its eight 128-thread groups then work through the brick lists exactly like the
// CTAs of brick_update_kernel.  Every node gather of the per-voxel tier becomes an LDS: ncu showed the global variant waiting on
// the L1-miss share of those gathers (long-scoreboard stalls = 52 % of all samples, L1 hit rate 74 %).
constexpr int UPDATE_SMEM_THREADS = 1024, UPDATE_SMEM_GROUPS = UPDATE_SMEM_THREADS / 128;
constexpr size_t UPDATE_SMEM_MAX_BYTES = 226 * 1024;   // of the 227 KB a CTA may own on sm_100
template <int KMAX, bool EXACTK, bool ONEVIEW>
__global__ void __launch_bounds__(UPDATE_SMEM_THREADS, 1) brick_update_smem_kernel(const __grid_constant__ ProjParams P, int n_nodes, int nbx, int nby,
                                                                                    int nbz, const uint8_t* cls, const uint32_t* stream_list,
                                                                                    const uint32_t* mixed_list, const float* region_rec) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4* s_rec = reinterpret_cast<float4*>(smem_raw);
    // behind the table: the open-voxel queues of the 32 warps
    uint32_t* wq = reinterpret_cast<uint32_t*>(smem_raw + (((size_t)n_nodes * DFB_NODE_REC_FLOATS * sizeof(float) + 127) & ~(size_t)127)) +
                   (size_t)(threadIdx.x >> 5) * 2 * WQ_CAP;
    WarpQueue Q = {wq, wq + WQ_CAP, 0};
    __shared__ __align__(8) unsigned long long bar;
    const uint32_t bytes = (uint32_t)n_nodes * (uint32_t)(DFB_NODE_REC_FLOATS * sizeof(float));
    const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
        const char* src = reinterpret_cast<const char*>(P.node_rec);
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(s_rec);
        for (uint32_t off = 0; off < bytes; off += 32768u) {
            const uint32_t n = bytes - off < 32768u ? bytes - off : 32768u;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + off),
                         "l"(src + off), "r"(n), "r"(bar_a)
                         : "memory");
        }
    }
    const uint32_t cnt_s = P.counters[2], cnt_m = P.counters[3];
    const uint32_t n_groups = gridDim.x * UPDATE_SMEM_GROUPS;
    const uint32_t n_t = cnt_m > n_groups ? cnt_m : n_groups;
    const uint32_t rec_a = (uint32_t)__cvta_generic_to_shared(s_rec);
    {   // wait for the table (phase 0 of the barrier)
        uint32_t done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar_a) : "memory");
    }
    // groups of one CTA take consecutive list entries in turn (neighbouring bricks: the same nodes, the same depth pixels)
    update_body<KMAX, EXACTK, ONEVIEW>(P, region_rec, nbx, nby, nbz, cls, stream_list, cnt_s, mixed_list, cnt_m,
                                       blockIdx.x * UPDATE_SMEM_GROUPS + (threadIdx.x >> 7), n_groups, n_t, threadIdx.x & 127, RecShared{rec_a}, Q);
}

// One deferred voxel through the reference-exact tier (a2 / a3).
template <int KMAX, int KT>
__device__ __forceinline__ void exact_voxel(const ProjParams& P, size_t i, size_t nvox, size_t plane, const uint16_t* ids, float v, float w) {
    int xs, y, z;
    if (nvox <= 0xffffffffull) {   // 32-bit index arithmetic (the 64-bit divisions were 5 % of the kernel's instructions)
        const uint32_t i32 = (uint32_t)i, plane32 = (uint32_t)plane, rz32 = (uint32_t)P.rz;
        const uint32_t xq = i32 / plane32, rem = i32 - xq * plane32, yq = rem / rz32;
        xs = (int)xq; y = (int)yq; z = (int)(rem - yq * rz32);
    } else {
        xs = (int)(i / plane);
        const size_t rem = i - (size_t)xs * plane;
        y = (int)(rem / P.rz);
        z = (int)(rem - (size_t)y * P.rz);
    }
    int m, f;
    voxel_projective_exact<KT>(P, xs + P.x0, y, z, ids, &v, &w, &m, &f);
    if (m) {
        P.tsdf[i] = v;
        P.weight[i] = w;
    }
    if (P.mask_out) P.mask_out[i] = (uint8_t)m;
    if (P.frustum_out) P.frustum_out[i] = (uint8_t)f;
}

// mode: 0 = work list (+ overflow bitmap; without one, a re-scan of the volume when the list overflowed), 1 = every voxel
#ifndef DFB_EXACT_MINB
#define DFB_EXACT_MINB 8
#endif
template <int KMAX, int KT>
__global__ void __launch_bounds__(128, DFB_EXACT_MINB) proj_exact_kernel(const __grid_constant__ ProjParams P, int all_mode) {
    const size_t nvox = (size_t)(P.x1 - P.x0) * P.ry * P.rz;
    const uint32_t count = P.counters[0];
    const bool overflow = !all_mode && count > P.capacity;
    const bool sweep_bits = overflow && P.overflow_bits != nullptr;
    const bool use_list = !all_mode && (!overflow || sweep_bits);
    const size_t n = use_list ? (size_t)(overflow ? P.capacity : count) : nvox;
    const size_t plane = (size_t)P.ry * P.rz;
    const int kk = KT > 0 ? KT : P.k;
    uint32_t done = 0;
    if (use_list && KT > 0 && !P.rigid && nvox <= 0xffffffffull) {
        // Work-list path, software-pipelined: a voxel's state hangs on a chain of dependent gathers (list entry -> kNN ids,
        // v, w -> node data) ahead of ~1000 instructions of arithmetic.  The list entry is fetched two iterations ahead and
        // ids / v / w one iteration ahead, so only the node and depth gathers (L1 / L2 hits) are left inside an iteration.
        const uint32_t stride = gridDim.x * blockDim.x, cnt = (uint32_t)n;
        uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
        if (t < cnt) {
            uint32_t i0 = P.list[t];
            uint32_t i1 = t + stride < cnt && t + stride >= t ? P.list[t + stride] : i0;
            uint16_t ids0[KMAX];
            load_ids<KMAX>(P.knn, i0, KMAX, ids0);
            float v0 = P.tsdf[i0], w0 = P.weight[i0];
            for (;;) {
                const uint32_t tn = t + stride, tnn = tn + stride;
                const bool has1 = tn < cnt && tn > t;
                const uint32_t i2 = (has1 && tnn < cnt && tnn > tn) ? P.list[tnn] : i1;
                uint16_t ids1[KMAX];
                load_ids<KMAX>(P.knn, i1, KMAX, ids1);
                const float v1 = P.tsdf[i1], w1 = P.weight[i1];
                exact_voxel<KMAX, KT>(P, i0, nvox, plane, ids0, v0, w0);
                ++done;
                if (!has1) break;
                t = tn; i0 = i1; i1 = i2; v0 = v1; w0 = w1;
#pragma unroll
                for (int j = 0; j < KMAX; ++j) ids0[j] = ids1[j];
            }
        }
    } else {
        for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
            const size_t i = use_list ? (size_t)P.list[t] : t;
            uint16_t ids[KMAX];
            if (!P.rigid) load_ids<KMAX>(P.knn, i, kk, ids);
            if (!use_list && !all_mode) {
                const size_t xq = i / plane, rem = i - xq * plane;
                int m0, f0;
                if (voxel_projective_classify<KMAX>(P, (int)xq + P.x0, (int)(rem / P.rz), (int)(rem % P.rz), ids, &m0, &f0) != CLS_UNCERTAIN) continue;
            }
            exact_voxel<KMAX, KT>(P, i, nvox, plane, ids, P.tsdf[i], P.weight[i]);
            ++done;
        }
    }
    if (sweep_bits) {
        // the voxels that did not fit the list: one thread per voxel over the bitmap (a warp = one word, read by all its lanes,
        // then cleared for the next call), so that a frame that overflowed by a lot still runs coalesced
        const size_t nwords = (nvox + 31) / 32;
        const size_t warp0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
        const int lane = threadIdx.x & 31;
        for (size_t wd = warp0; wd < nwords; wd += nwarps) {
            const uint32_t bits = P.overflow_bits[wd];
            if (!bits) continue;
            __syncwarp();
            if (lane == 0) P.overflow_bits[wd] = 0u;
            if ((bits >> lane) & 1u) {
                const size_t i = wd * 32 + (size_t)lane;
                uint16_t ids[KMAX];
                if (!P.rigid) load_ids<KMAX>(P.knn, i, kk, ids);
                exact_voxel<KMAX, KT>(P, i, nvox, plane, ids, P.tsdf[i], P.weight[i]);
                ++done;
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) done += __shfl_xor_sync(0xffffffffu, done, o);
    if ((threadIdx.x & 31) == 0 && done) atomicAdd(P.counters + 1, done);
}

// ------------------------------------------------------------------------------------------------
// a1
// ------------------------------------------------------------------------------------------------
#ifndef DFB_VOL_FAST_MINB
#define DFB_VOL_FAST_MINB 6   // latency-bound (ids -> node records -> live corners): 3 -> 0.76 ms at 256^3, 4 -> 0.63, 5 / 6 / 8 -> 0.54
#endif
template <int KMAX>
__global__ void __launch_bounds__(256, DFB_VOL_FAST_MINB) vol_fast_kernel(const __grid_constant__ VolParams P) {
    const int z = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int xs = blockIdx.z;
    const bool in = z < P.rz;
    const size_t i = ((size_t)xs * P.ry + y) * P.rz + (in ? z : 0);
    int cls = CLS_SKIP;
    float wi = 0.f;
    if (in) {
        uint16_t ids[KMAX];
        if (P.k > 0) load_ids<KMAX>(P.knn, i, P.k, ids);
        cls = voxel_volume_classify<KMAX>(P, xs + P.x0, y, z, ids, &wi);
    }
    push_uncertain(in && cls == CLS_UNCERTAIN, (uint32_t)i, P.list, P.capacity, P.counters, P.overflow_bits);
    if (!in || cls == CLS_UNCERTAIN) return;
    if (cls == CLS_CLAMP) {
        float v = P.tsdf[i], w = P.weight[i];
        if (P.k > 0) {
            const float wt = (w == 0.f) ? wi : w;
            v = (v * wt + fmul(P.tdist_f, wi)) / (wi + wt);
            w = fminf(wi + wt, P.wmax_f);
        } else {
            v = (v * w + P.tdist_f) / (1.0f + w);
            w = fminf(1.0f + w, P.wmax_f);
        }
        P.tsdf[i] = v;
        P.weight[i] = w;
    }
    if (P.mask_out) P.mask_out[i] = (cls == CLS_CLAMP) ? 1 : 0;
}

template <int KMAX, int KT>
__device__ __forceinline__ void vol_exact_voxel(const VolParams& P, size_t i, size_t plane) {
    int xs, y, z;
    if (plane * (size_t)(P.x1 - P.x0) <= 0xffffffffull) {   // 32-bit index arithmetic
        const uint32_t i32 = (uint32_t)i, plane32 = (uint32_t)plane, rz32 = (uint32_t)P.rz;
        const uint32_t xq = i32 / plane32, rem = i32 - xq * plane32, yq = rem / rz32;
        xs = (int)xq; y = (int)yq; z = (int)(rem - yq * rz32);
    } else {
        xs = (int)(i / plane);
        const size_t rem = i - (size_t)xs * plane;
        y = (int)(rem / P.rz);
        z = (int)(rem - (size_t)y * P.rz);
    }
    uint16_t ids[KMAX];
    if (P.k > 0) load_ids<KMAX>(P.knn, i, P.k, ids);
    float v = P.tsdf[i], w = P.weight[i];
    const bool upd = voxel_volume_exact<KT>(P, xs + P.x0, y, z, ids, &v, &w);
    if (upd) {
        P.tsdf[i] = v;
        P.weight[i] = w;
    }
    if (P.mask_out) P.mask_out[i] = upd ? 1 : 0;
}

#ifndef DFB_VOL_EXACT_MINB
#define DFB_VOL_EXACT_MINB 8   // measured, 256^3 with every voxel in the band: 4 -> 2.50 ms, 6 -> 2.19 ms, 8 -> 2.04 ms
#endif
template <int KMAX, int KT>
__global__ void __launch_bounds__(128, DFB_VOL_EXACT_MINB) vol_exact_kernel(const __grid_constant__ VolParams P, int all_mode) {
    const size_t nvox = (size_t)(P.x1 - P.x0) * P.ry * P.rz;
    const uint32_t count = P.counters[0];
    const bool overflow = !all_mode && count > P.capacity;
    // the voxels that did not fit the list are in the bitmap (swept one thread per voxel, see proj_exact_kernel: the reference's own
    // a1 call, test.py:110, puts EVERY voxel inside the band); without a bitmap the volume is re-classified
    const bool sweep_bits = overflow && P.overflow_bits != nullptr;
    const bool use_list = !all_mode && (!overflow || sweep_bits);
    const size_t n = use_list ? (size_t)(overflow ? P.capacity : count) : nvox;
    const size_t plane = (size_t)P.ry * P.rz;
    uint32_t done = 0;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
        const size_t i = use_list ? (size_t)P.list[t] : t;
        if (!use_list && !all_mode) {
            const int xs = (int)(i / plane);
            const size_t rem = i - (size_t)xs * plane;
            uint16_t ids[KMAX];
            if (P.k > 0) load_ids<KMAX>(P.knn, i, P.k, ids);
            float wi;
            if (voxel_volume_classify<KMAX>(P, xs + P.x0, (int)(rem / P.rz), (int)(rem % P.rz), ids, &wi) != CLS_UNCERTAIN) continue;
        }
        vol_exact_voxel<KMAX, KT>(P, i, plane);
        ++done;
    }
    if (sweep_bits) {
        const size_t nwords = (nvox + 31) / 32;
        const size_t warp0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
        const int lane = threadIdx.x & 31;
        for (size_t wd = warp0; wd < nwords; wd += nwarps) {
            const uint32_t bits = P.overflow_bits[wd];
            if (!bits) continue;
            __syncwarp();
            if (lane == 0) P.overflow_bits[wd] = 0u;
            if ((bits >> lane) & 1u) {
                vol_exact_voxel<KMAX, KT>(P, wd * 32 + (size_t)lane, plane);
                ++done;
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) done += __shfl_xor_sync(0xffffffffu, done, o);
    if ((threadIdx.x & 31) == 0 && done) atomicAdd(P.counters + 1, done);
}

// ------------------------------------------------------------------------------------------------
// a4: warp arbitrary points (exact tier)
// ------------------------------------------------------------------------------------------------
__global__ void warp_points_kernel(const float* pts, const float* normals, int64_t m, const int32_t* idx, int k,
                                   const float* node_pos, const float* node_dq, const float* node_w,
                                   const __grid_constant__ dfb_warpfield wf, double* out_p, double* out_n) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < m; t += (int64_t)gridDim.x * blockDim.x) {
        int ids[DFB_MAX_K];
        for (int j = 0; j < k; ++j) ids[j] = idx[t * k + j];
        double op[3], on[3];
        warp_ref(pts + 3 * t, normals ? normals + 3 * t : nullptr, ids, k, node_pos, node_dq, node_w, wf.lw, wf.has_lw != 0,
                 wf.lw_is_f32 != 0, op, on, nullptr);
        for (int c = 0; c < 3; ++c) out_p[3 * t + c] = op[c];
        if (normals && out_n)
            for (int c = 0; c < 3; ++c) out_n[3 * t + c] = on[c];
    }
}

__global__ void dq_blend_points_kernel(const float* pts, int64_t m, const int32_t* idx, int k, const float* node_pos,
                                       const float* node_dq, const float* node_w, double* out) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < m; t += (int64_t)gridDim.x * blockDim.x) {
        int ids[DFB_MAX_K];
        for (int j = 0; j < k; ++j) ids[j] = idx[t * k + j];
        double se3[8];
        dq_blend_ref(pts + 3 * t, ids, k, node_pos, node_dq, node_w, se3, nullptr);
        for (int c = 0; c < 8; ++c) out[8 * t + c] = se3[c];
    }
}

// ------------------------------------------------------------------------------------------------
// node packing
// ------------------------------------------------------------------------------------------------
__global__ void nodes_pack_kernel(const float* pos, const float* dq, const float* w, int n, float* rec) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float* r = rec + (size_t)i * DFB_NODE_REC_FLOATS;
    r[0] = pos[3 * i]; r[1] = pos[3 * i + 1]; r[2] = pos[3 * i + 2];
    const double ww = (double)w[i];
    r[3] = (float)(-1.4426950408889634 / (4.0 * ww * ww));
    for (int c = 0; c < 8; ++c) r[4 + c] = dq[8 * i + c];
}

dim3 fast_grid(const dfb_volume* vol, int& threads) {
    threads = vol->rz >= 256 ? 256 : ((vol->rz + 31) / 32) * 32;
    return dim3((vol->rz + threads - 1) / threads, vol->ry, vol->x1 - vol->x0);
}

int exact_blocks(size_t nvox) {
    static int per_sm = 0;
    // grid of the exact pass in CTAs per SM (8 are resident): 8 / 16 / 24 / 32 -> 0.160 / 0.153 / 0.150 / 0.149 ms at 512^3
    if (per_sm == 0) { const char* e = getenv("DFB_EXACT_CTAS_PER_SM"); per_sm = e ? atoi(e) : 32; if (per_sm < 1) per_sm = 32; }
    const size_t want = (nvox + 127) / 128, cap = (size_t)148 * per_sm;
    return (int)(want < cap ? (want ? want : 1) : cap);
}

struct BrickArgs {
    int n_nodes;
    const uint16_t* nodes;
    const uint8_t* count;
    const uint32_t* pairs;
    const uint16_t* rnodes;
    const uint8_t* rcount;
    const uint32_t* rpairs;
    float* rrec;
    uint8_t* cls;
    uint32_t* lists;
};

int run_projective(ProjParams& P, int mode, cudaStream_t s, const dfb_volume* vol, const BrickArgs& B) {
    const bool bricks = B.cls != nullptr && B.lists != nullptr && (P.rigid || (B.nodes != nullptr && B.count != nullptr));
    const bool do_classify = mode == DFB_MODE_HYBRID || mode == DFB_MODE_FAST_ONLY || mode == DFB_MODE_BRICK_CLASSIFY;
    const bool do_stream = mode == DFB_MODE_HYBRID || mode == DFB_MODE_FAST_ONLY || mode == DFB_MODE_BRICK_STREAM || mode == DFB_MODE_BRICK_UPDATE;
    const bool do_mixed = mode == DFB_MODE_HYBRID || mode == DFB_MODE_FAST_ONLY || mode == DFB_MODE_BRICK_MIXED || mode == DFB_MODE_BRICK_UPDATE;
    if (mode >= DFB_MODE_BRICK_CLASSIFY) DFB_REQUIRE(bricks, "brick modes need the brick workspace / candidate sets");
    // Overflow marks are consumed (and cleared) by the exact pass of the SAME call.  The profiling modes that stop before it must
    // not leave marks behind -- a later hybrid call would send those voxels through the exact tier on top of their fast-tier
    // update -- so they run without the bitmap (their deferred voxels beyond the list capacity are simply dropped).
    if (mode != DFB_MODE_HYBRID && mode != DFB_MODE_EXACT) P.overflow_bits = nullptr;
    if (mode == DFB_MODE_HYBRID || mode == DFB_MODE_EXACT || mode == DFB_MODE_FAST_ONLY || mode == DFB_MODE_BRICK_CLASSIFY)
        DFB_CUDA(cudaMemsetAsync(P.counters, 0, 8 * sizeof(uint32_t), s));
    const size_t nvox = (size_t)(P.x1 - P.x0) * P.ry * P.rz;
    if (mode != DFB_MODE_EXACT && mode != DFB_MODE_LIST_ONLY) {
        if (bricks) {
            const int nbx = (P.x1 - P.x0 + BRICK_X - 1) / BRICK_X, nby = (P.ry + BRICK_Y - 1) / BRICK_Y, nbz = (P.rz + BRICK_Z - 1) / BRICK_Z;
            const int nb = nbx * nby * nbz;
            DFB_REQUIRE(nbx <= BRICK_PACK_MAX_XY && nby <= BRICK_PACK_MAX_XY && nbz <= BRICK_PACK_MAX_Z,
                        "volume too large for the brick work lists (4096 x 4096 x 131072 voxels per slab)");
            const float* rrec = nullptr;
            const bool have_regions = !P.rigid && B.rnodes && B.rcount && B.rpairs && B.rrec;
            uint32_t* stream_list = B.lists;
            uint32_t* mixed_list = B.lists + nb;
            const bool classified = false;
            // (Measured and dropped: ONE kernel with a warp per region doing bound -> whole-region test -> its 16 bricks.  No block
            // barriers, one launch less -- but 0.245 instead of 0.151 ms at 512^3: the serial chain per warp is long and its ~130
            // registers leave two CTAs per SM.)
            if (have_regions) {
                rrec = B.rrec;
                if (do_classify) {
                    const int nrx = (P.x1 - P.x0 + REGION_X - 1) / REGION_X, nry = (P.ry + REGION_Y - 1) / REGION_Y, nrz = (P.rz + REGION_Z - 1) / REGION_Z;
                    region_bounds_kernel<<<dim3(nrz, nry, nrx), DFB_REGION_THREADS, 0, s>>>(P, B.rnodes, B.rcount, B.rpairs, nry, nrz, B.rrec);
                    DFB_LAUNCH_CHECK("region_bounds_kernel");
                }
            }
            static int upd_per_sm = 0;
            // grid of the update pass in CTAs per SM (8 are resident): flat between 16 and 64 (0.380 ms), 12 -> 0.433 ms
            if (upd_per_sm == 0) { const char* e = getenv("DFB_UPDATE_CTAS_PER_SM"); upd_per_sm = e ? atoi(e) : 16; if (upd_per_sm < 1) upd_per_sm = 16; }
            const int grid = nb < 148 * upd_per_sm ? nb : 148 * upd_per_sm;
            if (do_classify && !classified) {
                static int G = 0;
                if (G == 0) { const char* e = getenv("DFB_CLASSIFY_G"); G = e ? atoi(e) : 8; }
                const int per_cta = 128 / G;
                const int nb_pad = ((nbx + 3) / 4) * ((nby + 3) / 4) * nbz * 16;   // region-major enumeration (edge regions padded)
                const int cgrid = (nb_pad + per_cta - 1) / per_cta < 148 * 32 ? (nb_pad + per_cta - 1) / per_cta : 148 * 32;
                if (G == 1) brick_classify_kernel<1><<<cgrid, 128, 0, s>>>(P, B.nodes, B.count, B.pairs, rrec, nbx, nby, nbz, B.cls, stream_list, mixed_list);
                else if (G == 2) brick_classify_kernel<2><<<cgrid, 128, 0, s>>>(P, B.nodes, B.count, B.pairs, rrec, nbx, nby, nbz, B.cls, stream_list, mixed_list);
                else if (G == 4) brick_classify_kernel<4><<<cgrid, 128, 0, s>>>(P, B.nodes, B.count, B.pairs, rrec, nbx, nby, nbz, B.cls, stream_list, mixed_list);
                else if (G == 16) brick_classify_kernel<16><<<cgrid, 128, 0, s>>>(P, B.nodes, B.count, B.pairs, rrec, nbx, nby, nbz, B.cls, stream_list, mixed_list);
                else if (G == 32) brick_classify_kernel<32><<<cgrid, 128, 0, s>>>(P, B.nodes, B.count, B.pairs, rrec, nbx, nby, nbz, B.cls, stream_list, mixed_list);
                else brick_classify_kernel<8><<<cgrid, 128, 0, s>>>(P, B.nodes, B.count, B.pairs, rrec, nbx, nby, nbz, B.cls, stream_list, mixed_list);
                DFB_LAUNCH_CHECK("brick_classify_kernel");
            }
            if (do_stream && do_mixed) {
#define DFB_BRICK_DISPATCH(KERNEL, ...)                                                                                  \
    do {                                                                                                                 \
        const bool one = P.n_views == 1;                                                                                 \
        if (P.k == 4 && !P.rigid) { if (one) KERNEL<4, true, true><<<grid, 128, 0, s>>>(__VA_ARGS__); else KERNEL<4, true, false><<<grid, 128, 0, s>>>(__VA_ARGS__); } \
        else if (P.k == 8 && !P.rigid) { if (one) KERNEL<8, true, true><<<grid, 128, 0, s>>>(__VA_ARGS__); else KERNEL<8, true, false><<<grid, 128, 0, s>>>(__VA_ARGS__); } \
        else if (P.k <= 4) { if (one) KERNEL<4, false, true><<<grid, 128, 0, s>>>(__VA_ARGS__); else KERNEL<4, false, false><<<grid, 128, 0, s>>>(__VA_ARGS__); } \
        else { if (one) KERNEL<8, false, true><<<grid, 128, 0, s>>>(__VA_ARGS__); else KERNEL<8, false, false><<<grid, 128, 0, s>>>(__VA_ARGS__); } \
    } while (0)
                // Shared-memory node table (brick_update_smem_kernel) or node gathers through L1 (brick_update_kernel).  Measured
                // (profiles/r2_ab_update.md): one view, k = 4 at 512^3 -- 0.397 ms with the table, 0.377 ms without (the quad pre-test
                // leaves a quarter of the MIXED voxels to gather nodes at all, and eight 128-thread CTAs per SM balance better than one
                // CTA of 1024); eight views, k = 8 at 256^3 -- 0.61 ms with the table, 0.69 ms without.  DFB_UPDATE_SMEM=0/1 forces one.
                static int smem_env = -2;
                if (smem_env == -2) { const char* e = getenv("DFB_UPDATE_SMEM"); smem_env = e ? atoi(e) : -1; }
                const int use_smem = smem_env >= 0 ? smem_env : (P.n_views > 1 || P.k > 4);
                static int use_quads = -1;
                if (use_quads < 0) { const char* e = getenv("DFB_QUADS"); use_quads = e ? atoi(e) : 1; }
                const float* qrec = use_quads ? rrec : nullptr;
                const size_t rec_bytes = (((size_t)B.n_nodes * DFB_NODE_REC_FLOATS * sizeof(float) + 127) & ~(size_t)127) + (UPDATE_SMEM_THREADS / 32) * WQ_BYTES;
                if (use_smem && !P.rigid && B.n_nodes > 0 && rec_bytes <= UPDATE_SMEM_MAX_BYTES) {
                    // one persistent CTA per SM, the node table in its shared memory
                    static int n_sm = 0;
                    if (n_sm == 0) {
                        int dev = 0;
                        DFB_CUDA(cudaGetDevice(&dev));
                        DFB_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
                    }
                    const int sgrid = nb < n_sm * UPDATE_SMEM_GROUPS ? (nb + UPDATE_SMEM_GROUPS - 1) / UPDATE_SMEM_GROUPS : n_sm;
#define DFB_SMEM_LAUNCH(K_, E_, O_)                                                                                              \
    do {                                                                                                                         \
        DFB_CUDA(cudaFuncSetAttribute(brick_update_smem_kernel<K_, E_, O_>, cudaFuncAttributeMaxDynamicSharedMemorySize,        \
                                      (int)UPDATE_SMEM_MAX_BYTES));                                                              \
        brick_update_smem_kernel<K_, E_, O_><<<sgrid, UPDATE_SMEM_THREADS, rec_bytes, s>>>(P, B.n_nodes, nbx, nby, nbz, B.cls,   \
                                                                                          stream_list, mixed_list, qrec);      \
    } while (0)
                    const bool one = P.n_views == 1;
                    if (P.k == 4) { if (one) DFB_SMEM_LAUNCH(4, true, true); else DFB_SMEM_LAUNCH(4, true, false); }
                    else if (P.k == 8) { if (one) DFB_SMEM_LAUNCH(8, true, true); else DFB_SMEM_LAUNCH(8, true, false); }
                    else if (P.k < 4) { if (one) DFB_SMEM_LAUNCH(4, false, true); else DFB_SMEM_LAUNCH(4, false, false); }
                    else { if (one) DFB_SMEM_LAUNCH(8, false, true); else DFB_SMEM_LAUNCH(8, false, false); }
                    DFB_LAUNCH_CHECK("brick_update_smem_kernel");
                } else {
                    DFB_BRICK_DISPATCH(brick_update_kernel, P, nbx, nby, nbz, B.cls, stream_list, mixed_list, qrec);
                    DFB_LAUNCH_CHECK("brick_update_kernel");
                }
            } else if (do_stream) {
                brick_stream_kernel<<<grid, 128, 0, s>>>(P, nbx, nby, nbz, B.cls, stream_list);
                DFB_LAUNCH_CHECK("brick_stream_kernel");
            } else if (do_mixed) {
                static int use_quads_m = -1;
                if (use_quads_m < 0) { const char* e = getenv("DFB_QUADS"); use_quads_m = e ? atoi(e) : 1; }
                DFB_BRICK_DISPATCH(brick_mixed_kernel, P, nbx, nby, nbz, B.cls, mixed_list, use_quads_m ? rrec : nullptr);
                DFB_LAUNCH_CHECK("brick_mixed_kernel");
            }
        } else {
            int threads;
            const dim3 grid = fast_grid(vol, threads);
            if (P.k <= 4) proj_fast_kernel<4><<<grid, threads, 0, s>>>(P);
            else proj_fast_kernel<8><<<grid, threads, 0, s>>>(P);
            DFB_LAUNCH_CHECK("proj_fast_kernel");
        }
        if (mode != DFB_MODE_HYBRID) return DFB_OK;
    }
    const int all = mode == DFB_MODE_EXACT ? 1 : 0;
    if (P.k == 4 && !P.rigid) proj_exact_kernel<4, 4><<<exact_blocks(nvox), 128, 0, s>>>(P, all);
    else if (P.k == 8 && !P.rigid) proj_exact_kernel<8, 8><<<exact_blocks(nvox), 128, 0, s>>>(P, all);
    else if (P.k <= 4) proj_exact_kernel<4, 0><<<exact_blocks(nvox), 128, 0, s>>>(P, all);
    else proj_exact_kernel<8, 0><<<exact_blocks(nvox), 128, 0, s>>>(P, all);
    DFB_LAUNCH_CHECK("proj_exact_kernel");
    return DFB_OK;
}

}  // namespace

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" int dfb_nodes_pack(const float* node_pos, const float* node_dq, const float* node_w, int n_nodes,
                              float* node_rec, dfb_stream_t stream) {
    DFB_REQUIRE(node_pos && node_dq && node_w && node_rec, "null pointer");
    DFB_REQUIRE(n_nodes > 0, "n_nodes must be positive");
    nodes_pack_kernel<<<(n_nodes + 127) / 128, 128, 0, (cudaStream_t)stream>>>(node_pos, node_dq, node_w, n_nodes, node_rec);
    DFB_LAUNCH_CHECK("nodes_pack_kernel");
    return DFB_OK;
}

extern "C" int dfb_tsdf_update_projective(const dfb_volume* vol, const dfb_warpfield* wf, const dfb_views* views,
                                          double tdist, double wmax, int mode, const dfb_workspace* ws,
                                          uint8_t* mask_out, uint8_t* frustum_out, dfb_stream_t stream) {
    ProjParams P;
    if (int r = build_projective(P, vol, wf, views, tdist, wmax, mode, ws, mask_out, frustum_out)) return r;
    const BrickArgs B = {wf->n_nodes, wf->brick_nodes, wf->brick_count, wf->brick_pairs, wf->region_nodes, wf->region_count, wf->region_pairs, wf->region_rec,
                         ws->brick_cls, ws->brick_lists};
    return run_projective(P, mode, (cudaStream_t)stream, vol, B);
}

extern "C" int dfb_fuse_depth_rigid(const dfb_volume* vol, int tsdf_res, const float* depth, int rows, int cols,
                                    const double lw34[12], const double K[9], const double Kinv[9], double scale,
                                    const double center[3], double tdist, double wmax, int mode,
                                    const dfb_workspace* ws, uint8_t* mask_out, uint8_t* frustum_out,
                                    dfb_stream_t stream) {
    ProjParams P;
    if (int r = build_rigid(P, vol, tsdf_res, depth, rows, cols, lw34, K, Kinv, scale, center, tdist, wmax, mode, ws,
                            mask_out, frustum_out))
        return r;
    const BrickArgs B = {0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, ws->brick_cls, ws->brick_lists};
    return run_projective(P, mode, (cudaStream_t)stream, vol, B);
}

extern "C" int dfb_tsdf_update_volume(const dfb_volume* vol, const dfb_warpfield* wf, const float* curr, int cx,
                                      int cy, int cz, double tdist, double wmax, int mode, const dfb_workspace* ws,
                                      uint8_t* mask_out, dfb_stream_t stream) {
    VolParams P;
    if (int r = build_volume(P, vol, wf, curr, cx, cy, cz, tdist, wmax, mode, ws, mask_out)) return r;
    cudaStream_t s = (cudaStream_t)stream;
    if (mode != DFB_MODE_HYBRID && mode != DFB_MODE_EXACT) P.overflow_bits = nullptr;   // see run_projective
    if (mode != DFB_MODE_LIST_ONLY) DFB_CUDA(cudaMemsetAsync(P.counters, 0, 8 * sizeof(uint32_t), s));
    const size_t nvox = (size_t)(P.x1 - P.x0) * P.ry * P.rz;
    if (mode == DFB_MODE_HYBRID || mode == DFB_MODE_FAST_ONLY) {
        int threads;
        const dim3 grid = fast_grid(vol, threads);
        if (P.k <= 4) vol_fast_kernel<4><<<grid, threads, 0, s>>>(P);
        else vol_fast_kernel<8><<<grid, threads, 0, s>>>(P);
        DFB_LAUNCH_CHECK("vol_fast_kernel");
        if (mode == DFB_MODE_FAST_ONLY) return DFB_OK;
    }
    const int all = mode == DFB_MODE_EXACT ? 1 : 0;
    if (P.k == 4) vol_exact_kernel<4, 4><<<exact_blocks(nvox), 128, 0, s>>>(P, all);
    else if (P.k == 8) vol_exact_kernel<8, 8><<<exact_blocks(nvox), 128, 0, s>>>(P, all);
    else if (P.k <= 4) vol_exact_kernel<4, 0><<<exact_blocks(nvox), 128, 0, s>>>(P, all);
    else vol_exact_kernel<8, 0><<<exact_blocks(nvox), 128, 0, s>>>(P, all);
    DFB_LAUNCH_CHECK("vol_exact_kernel");
    return DFB_OK;
}

extern "C" int dfb_warp_points(const float* pts, const float* normals, int64_t m, const int32_t* idx,
                               const dfb_warpfield* wf, double* out_pts, double* out_normals, dfb_stream_t stream) {
    if (int r = validate_warpfield(wf, false)) return r;
    DFB_REQUIRE(pts && out_pts && m >= 0, "null pointer / negative count");
    DFB_REQUIRE(wf->k == 0 || idx, "neighbour indices are null");
    if (m == 0) return DFB_OK;
    const int blocks = (int)((m + 127) / 128 < 148 * 8 ? (m + 127) / 128 : 148 * 8);
    warp_points_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>(pts, normals, m, idx, wf->k, wf->node_pos, wf->node_dq,
                                                                 wf->node_w, *wf, out_pts, out_normals);
    DFB_LAUNCH_CHECK("warp_points_kernel");
    return DFB_OK;
}

extern "C" int dfb_dq_blend_points(const float* pts, int64_t m, const int32_t* idx, const dfb_warpfield* wf, double* out_dq,
                                   dfb_stream_t stream) {
    if (int r = validate_warpfield(wf, false)) return r;
    DFB_REQUIRE(pts && idx && out_dq && m >= 0 && wf->k > 0, "bad arguments");
    if (m == 0) return DFB_OK;
    const int blocks = (int)((m + 127) / 128 < 148 * 8 ? (m + 127) / 128 : 148 * 8);
    dq_blend_points_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>(pts, m, idx, wf->k, wf->node_pos, wf->node_dq, wf->node_w, out_dq);
    DFB_LAUNCH_CHECK("dq_blend_points_kernel");
    return DFB_OK;
}

extern "C" int64_t dfb_brick_count(int sx, int ry, int rz) {
    return (int64_t)((sx + BRICK_X - 1) / BRICK_X) * ((ry + BRICK_Y - 1) / BRICK_Y) * ((rz + BRICK_Z - 1) / BRICK_Z);
}

extern "C" int dfb_brick_nodes_build(const uint16_t* knn, int k, int rx, int ry, int rz, int x0, int x1, uint16_t* brick_nodes,
                                     uint8_t* brick_count, uint32_t* brick_pairs, dfb_stream_t stream) {
    return dfb_brick_nodes_update(knn, k, rx, ry, rz, x0, x1, nullptr, brick_nodes, brick_count, brick_pairs, stream);
}

extern "C" int dfb_brick_nodes_update(const uint16_t* knn, int k, int rx, int ry, int rz, int x0, int x1, const uint8_t* dirty8, uint16_t* brick_nodes,
                                      uint8_t* brick_count, uint32_t* brick_pairs, dfb_stream_t stream) {
    DFB_REQUIRE(knn && brick_nodes && brick_count && brick_pairs, "null pointer");
    DFB_REQUIRE(k >= 1 && k <= DFB_MAX_K && rx > 0 && ry > 0 && rz > 0 && x0 >= 0 && x1 > x0 && x1 <= rx, "bad arguments");
    const int sx = x1 - x0;
    const int nby = (ry + BRICK_Y - 1) / BRICK_Y, nbz = (rz + BRICK_Z - 1) / BRICK_Z;
    const int64_t nb = dfb_brick_count(sx, ry, rz);
    DFB_REQUIRE(nb < ((int64_t)1 << 31), "too many bricks");
    brick_nodes_kernel<<<(unsigned)nb, 128, 0, (cudaStream_t)stream>>>(knn, k, sx, ry, rz, nby, nbz, brick_nodes, brick_count, brick_pairs, dirty8);
    DFB_LAUNCH_CHECK("brick_nodes_kernel");
    return DFB_OK;
}

extern "C" int64_t dfb_region_count(int sx, int ry, int rz) {
    return (int64_t)((sx + REGION_X - 1) / REGION_X) * ((ry + REGION_Y - 1) / REGION_Y) * ((rz + REGION_Z - 1) / REGION_Z);
}

extern "C" int dfb_region_build(const uint16_t* knn, int k, int rx, int ry, int rz, int x0, int x1, uint16_t* region_nodes,
                                uint8_t* region_count, uint32_t* region_pairs, dfb_stream_t stream) {
    return dfb_region_update(knn, k, rx, ry, rz, x0, x1, nullptr, region_nodes, region_count, region_pairs, stream);
}

extern "C" int dfb_region_update(const uint16_t* knn, int k, int rx, int ry, int rz, int x0, int x1, const uint8_t* dirty8, uint16_t* region_nodes,
                                 uint8_t* region_count, uint32_t* region_pairs, dfb_stream_t stream) {
    DFB_REQUIRE(knn && region_nodes && region_count && region_pairs, "null pointer");
    DFB_REQUIRE(k >= 1 && k <= DFB_MAX_K && rx > 0 && ry > 0 && rz > 0 && x0 >= 0 && x1 > x0 && x1 <= rx, "bad arguments");
    const int sx = x1 - x0;
    const int nry = (ry + REGION_Y - 1) / REGION_Y, nrz = (rz + REGION_Z - 1) / REGION_Z;
    const int64_t nr = dfb_region_count(sx, ry, rz);
    DFB_REQUIRE(nr < ((int64_t)1 << 31), "too many regions");
    region_build_kernel<<<(unsigned)nr, 256, 0, (cudaStream_t)stream>>>(knn, k, sx, ry, rz, nry, nrz, region_nodes, region_count, region_pairs, dirty8);
    DFB_LAUNCH_CHECK("region_build_kernel");
    return DFB_OK;
}
