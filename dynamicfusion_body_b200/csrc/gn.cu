// gn.cu -- warp-field Gauss-Newton path: residuals (reference arithmetic), analytic Jacobians, block-sparse
// normal equations J^T W J / J^T W f, damped node-space solve (block-Jacobi PCG) and node update.
//
// The reference hands `computef` to scipy.optimize.least_squares(jac='2-point') (core/fusion.py:382-392): the
// Jacobian is finite-differenced (93.9 % of its solve time) and J^T J is never formed.  Here every data residual
// contributes a rank-1 8x8 outer product g g^T scaled by the blend weights w_a w_b to the k x k node blocks it
// touches (d r / d dq_a = w_a * g), so J^T J is assembled directly into a BSR matrix over the node co-occurrence
// graph.  It is a scatter-add of tiny outer products, not a dense contraction: CUDA cores + float64 atomics,
// no tensor cores.  Everything is float64 so that the solved update agrees with the float64
// oracle to ~1e-9, far inside the 1e-4 tolerance of the north star.
#include <cooperative_groups.h>

#include "common.h"
#include "dfb_gn.h"

namespace cg = cooperative_groups;

using namespace dfb;

namespace {

GNParams to_params(const dfb_gn_problem* p) {
    GNParams P;
    P.n_vert = p->n_vert; P.vertices = p->vertices; P.normals = p->normals; P.corr = p->corr; P.vert_knn = p->vert_knn;
    P.order = p->order;
    P.n_nodes = p->n_nodes; P.k = p->k; P.node_pos = p->node_pos; P.node_w = p->node_w; P.node_nbr = p->node_nbr;
    for (int i = 0; i < 8; ++i) P.lw[i] = p->lw[i];
    P.lw_is_f32 = p->lw_is_f32;
    dq_to_affine(p->lw, P.A);
    P.rw = p->rw; P.huber = p->huber; P.f_scale = p->f_scale > 0 ? p->f_scale : 1.0;
    return P;
}

int validate(const dfb_gn_problem* p) {
    DFB_REQUIRE(p, "problem is null");
    DFB_REQUIRE(p->n_vert >= 0 && p->n_nodes > 0 && p->k >= 1 && p->k <= DFB_MAX_K, "bad problem sizes");
    DFB_REQUIRE(p->n_nodes >= p->k, "n_nodes < k");
    DFB_REQUIRE(p->node_pos && p->node_w && p->node_nbr, "node arrays are null");
    DFB_REQUIRE(p->n_vert == 0 || (p->vertices && p->normals && p->corr && p->vert_knn), "vertex arrays are null");
    return DFB_OK;
}

// ---- residual values (a9 / a11) -----------------------------------------------------------------------------
__global__ void residual_kernel(const __grid_constant__ GNParams P, const double* x, int x_is_f32, double* f) {
    const int64_t n_reg = (int64_t)P.n_nodes * P.k;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < P.n_vert + n_reg; t += (int64_t)gridDim.x * blockDim.x) {
        if (t < P.n_vert) {
            f[t] = data_residual_ref(P, x, x_is_f32 != 0, P.lw, P.lw_is_f32 != 0, t);
        } else {
            const int64_t q = t - P.n_vert;
            double r[3];
            reg_residual_ref(P, x, x_is_f32 != 0, (int)(q / P.k), (int)(q % P.k), r);
            for (int c = 0; c < 3; ++c) f[P.n_vert + 3 * q + c] = r[c];
        }
    }
}

struct LwArg { double lw[8]; int is_f32; };

__global__ void residual_lw_kernel(const __grid_constant__ GNParams P, const double* dq, int dq_is_f32, const __grid_constant__ LwArg L, double* f) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < P.n_vert; t += (int64_t)gridDim.x * blockDim.x)
        f[t] = data_residual_ref(P, dq, dq_is_f32 != 0, L.lw, L.is_f32 != 0, t);
}

// ---- sparsity pattern of J^T J over node pairs ----------------------------------------------------------------
__device__ __forceinline__ void mark(uint32_t* bitmap, int words, int a, int b) {
    atomicOr(bitmap + (size_t)a * words + (b >> 5), 1u << (b & 31));
}

__global__ void pattern_mark_kernel(const __grid_constant__ GNParams P, uint32_t* bitmap, int words) {
    const int64_t n_reg = (int64_t)P.n_nodes * P.k;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < P.n_vert + n_reg + P.n_nodes; t += (int64_t)gridDim.x * blockDim.x) {
        if (t < P.n_vert) {
            const int32_t* ids = P.vert_knn + t * P.k;
            for (int a = 0; a < P.k; ++a)
                for (int b = 0; b < P.k; ++b) mark(bitmap, words, ids[a], ids[b]);
        } else if (t < P.n_vert + n_reg) {
            const int64_t q = t - P.n_vert;
            const int i = (int)(q / P.k), j = P.node_nbr[q];
            mark(bitmap, words, i, i); mark(bitmap, words, i, j); mark(bitmap, words, j, i); mark(bitmap, words, j, j);
        } else {
            const int i = (int)(t - P.n_vert - n_reg);
            mark(bitmap, words, i, i);  // the damping term needs every diagonal block
        }
    }
}

__global__ void pattern_count_kernel(const uint32_t* bitmap, int n, int words, int32_t* row_ptr) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int c = 0;
    for (int w = 0; w < words; ++w) c += __popc(bitmap[(size_t)i * words + w]);
    row_ptr[i + 1] = c;
    if (i == 0) row_ptr[0] = 0;
}

__global__ void pattern_scan_kernel(int32_t* row_ptr, int n) {  // single block, in-place inclusive scan of row_ptr[1..n]
    __shared__ int32_t carry;
    __shared__ int32_t buf[1024];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        buf[threadIdx.x] = (i < n) ? row_ptr[i + 1] : 0;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            const int32_t v = (threadIdx.x >= o) ? buf[threadIdx.x - o] : 0;
            __syncthreads();
            buf[threadIdx.x] += v;
            __syncthreads();
        }
        if (i < n) row_ptr[i + 1] = buf[threadIdx.x] + carry;
        __syncthreads();
        if (threadIdx.x == 1023) carry += buf[1023];
        __syncthreads();
    }
}

__global__ void pattern_fill_kernel(const uint32_t* bitmap, int n, int words, const int32_t* row_ptr, int32_t* col_idx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int o = row_ptr[i];
    for (int w = 0; w < words; ++w) {
        uint32_t m = bitmap[(size_t)i * words + w];
        while (m) {
            const int b = __ffs(m) - 1;
            col_idx[o++] = w * 32 + b;
            m &= m - 1;
        }
    }
}

__device__ __forceinline__ int find_slot(const int32_t* row_ptr, const int32_t* col_idx, int a, int b) {
    int lo = row_ptr[a], hi = row_ptr[a + 1] - 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (col_idx[mid] < b) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// ---- normal equations -------------------------------------------------------------------------------------------
// data term.  d r / d dq_a = w_a g with ONE 8-vector g per residual, so a residual adds (om w_a w_b) g g^T to block (a, b): residuals
// that share their node tuple add to the SAME k(k+1)/2 blocks.  The host hands over a processing order sorted by node tuple
// (dfb_gn_problem.order); each lane evaluates the (r, g, wts) of its OWN residual, then the warp walks its 32 residuals, broadcasting
// one residual's values by shuffle, every lane owning two of the 64 entries of each 8x8 block -- and ACCUMULATES IN REGISTERS for as
// long as the node tuple stays the same.  Only a change of tuple (and the end of the 32) flushes: k(k+1)/2 coalesced 512-byte atomic
// bursts + the block look-ups, instead of once per residual (300 k residuals over ~1 k nodes: ~15x fewer atomics).
// Only the blocks of the upper triangle of the node graph are accumulated; mirror_lower_kernel copies them.
template <int K>
__global__ void __launch_bounds__(256) normal_eq_data_kernel(const __grid_constant__ GNParams P, const double* x, const int32_t* row_ptr,
                                                            const int32_t* col_idx, double* H, double* g, double* cost) {
    constexpr int NB = K * (K + 1) / 2;
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int r0 = lane >> 3, c0 = lane & 7;
    double c_rob = 0.0, c_l2 = 0.0;
    // the warp's batches are CONSECUTIVE in the processing order, so a run of one tuple is cut by as few batch ends as possible
    const int64_t n_batches = (P.n_vert + 31) / 32;
    const int64_t per_warp = (n_batches + nwarps - 1) / nwarps;
    const int64_t b_first = warp * per_warp, b_last = b_first + per_warp < n_batches ? b_first + per_warp : n_batches;
    int run_ids[K];
    double acc0[NB], acc1[NB], gacc[K];
    bool have_run = false;
#pragma unroll
    for (int a = 0; a < K; ++a) { run_ids[a] = -1; gacc[a] = 0.0; }
#pragma unroll
    for (int t = 0; t < NB; ++t) { acc0[t] = 0.0; acc1[t] = 0.0; }
    auto flush = [&]() {
        int t = 0;
#pragma unroll
        for (int a = 0; a < K; ++a) {
#pragma unroll
            for (int b = a; b < K; ++b, ++t) {
                const int ia = run_ids[a], ib = run_ids[b];               // ascending: ia <= ib
                const int slot = find_slot(row_ptr, col_idx, ia, ib);
                atomicAdd(H + (size_t)slot * 64 + lane, acc0[t]);
                atomicAdd(H + (size_t)slot * 64 + 32 + lane, acc1[t]);
                acc0[t] = 0.0; acc1[t] = 0.0;
            }
            if (lane < 8) atomicAdd(g + 8 * (size_t)run_ids[a] + lane, gacc[a]);
            gacc[a] = 0.0;
        }
    };
    for (int64_t batch = b_first; batch < b_last; ++batch) {
        const int64_t base = batch * 32;
        const int64_t slot_i = base + lane;
        double r = 0.0, gv[8], wts[K], om = 0.0;
        int ids[K];
#pragma unroll
        for (int t = 0; t < 8; ++t) gv[t] = 0.0;
#pragma unroll
        for (int a = 0; a < K; ++a) { wts[a] = 0.0; ids[a] = 0; }
        if (slot_i < P.n_vert) {
            const int64_t mine = P.order ? (int64_t)P.order[slot_i] : slot_i;
            double wfull[DFB_MAX_K];
            data_residual_jac(P, x, mine, &r, gv, wfull);
            om = huber_weight(r, P.huber, P.f_scale);
            c_rob += huber_rho(r, P.huber, P.f_scale);
            c_l2 += 0.5 * r * r;
#pragma unroll
            for (int a = 0; a < K; ++a) { ids[a] = P.vert_knn[mine * K + a]; wts[a] = wfull[a]; }
            // canonical node order within the residual (ascending id): block (a, b) only depends on (id_a, id_b, w_a w_b)
#pragma unroll
            for (int a = 1; a < K; ++a)
#pragma unroll
                for (int b = a; b > 0; --b)
                    if (ids[b] < ids[b - 1]) {
                        const int ti = ids[b]; ids[b] = ids[b - 1]; ids[b - 1] = ti;
                        const double tw = wts[b]; wts[b] = wts[b - 1]; wts[b - 1] = tw;
                    }
        }
        const int cnt = (int)((P.n_vert - base < 32) ? (P.n_vert - base) : 32);
        for (int src = 0; src < cnt; ++src) {
            int sid[K];
            bool same = have_run;
#pragma unroll
            for (int a = 0; a < K; ++a) {
                sid[a] = __shfl_sync(0xffffffffu, ids[a], src);
                same = same && sid[a] == run_ids[a];
            }
            if (!same) {                                                   // warp-uniform
                if (have_run) flush();
#pragma unroll
                for (int a = 0; a < K; ++a) run_ids[a] = sid[a];
                have_run = true;
            }
            const double s_r = __shfl_sync(0xffffffffu, r, src), s_om = __shfl_sync(0xffffffffu, om, src);
            double sg[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) sg[t] = __shfl_sync(0xffffffffu, gv[t], src);
            // this lane's two entries of g g^T: (r0, c0) and (r0 + 4, c0); select without dynamic register indexing
            double ga = 0.0, gb = 0.0, gc = 0.0, gl = 0.0;
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                if (t == r0) ga = sg[t];
                if (t == r0 + 4) gb = sg[t];
                if (t == c0) gc = sg[t];
                if (t == (lane & 7)) gl = sg[t];
            }
            const double e0 = s_om * ga * gc, e1 = s_om * gb * gc;
            double sw[K];
#pragma unroll
            for (int a = 0; a < K; ++a) sw[a] = __shfl_sync(0xffffffffu, wts[a], src);
            int t = 0;
#pragma unroll
            for (int a = 0; a < K; ++a) {
#pragma unroll
                for (int b = a; b < K; ++b, ++t) {
                    // a node listed twice in one residual (never from a kNN query): blocks (a, b) and (b, a) coincide
                    const double cf = sw[a] * sw[b] * ((b != a && sid[a] == sid[b]) ? 2.0 : 1.0);
                    acc0[t] += cf * e0;
                    acc1[t] += cf * e1;
                }
                gacc[a] += s_om * sw[a] * s_r * gl;
            }
        }
    }
    if (have_run) flush();
    for (int o = 16; o > 0; o >>= 1) { c_rob += __shfl_xor_sync(0xffffffffu, c_rob, o); c_l2 += __shfl_xor_sync(0xffffffffu, c_l2, o); }
    if (lane == 0 && (c_rob != 0.0 || c_l2 != 0.0)) { atomicAdd(cost, c_rob); atomicAdd(cost + 1, c_l2); }
}

// lower-triangle blocks of the data term = their upper-triangle twins (see normal_eq_data_kernel); 64 threads per block
__global__ void mirror_lower_kernel(const int32_t* row_ptr, const int32_t* col_idx, int n, double* H) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t s = t >> 6;
    const int e = (int)(t & 63);
    if (s >= row_ptr[n]) return;
    const int j = col_idx[s];
    // row of slot s: largest i with row_ptr[i] <= s
    int lo = 0, hi = n;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (row_ptr[mid] <= s) lo = mid; else hi = mid;
    }
    const int i = lo;
    if (j >= i) return;
    H[(size_t)s * 64 + e] = H[(size_t)find_slot(row_ptr, col_idx, j, i) * 64 + e];
}

__device__ __forceinline__ void add_block(double* Hb, const double A[3][8], const double B[3][8], const double* om) {
    for (int r = 0; r < 8; ++r)
        for (int c = 0; c < 8; ++c) {
            const double v = om[0] * A[0][r] * B[0][c] + om[1] * A[1][r] * B[1][c] + om[2] * A[2][r] * B[2][c];
            if (v != 0.0) atomicAdd(Hb + r * 8 + c, v);
        }
}

__global__ void normal_eq_reg_kernel(const __grid_constant__ GNParams P, const double* x, const int32_t* row_ptr, const int32_t* col_idx,
                                     double* H, double* g, double* cost) {
    const int64_t n_reg = (int64_t)P.n_nodes * P.k;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n_reg; q += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(q / P.k), jj = (int)(q % P.k);
        double r[3], Ji[3][8], Jj[3][8], om[3];
        const int j = reg_residual_jac(P, x, i, jj, r, Ji, Jj);
        double c_rob = 0.0, c_l2 = 0.0;
        for (int t = 0; t < 3; ++t) {
            om[t] = huber_weight(r[t], P.huber, P.f_scale);
            c_rob += huber_rho(r[t], P.huber, P.f_scale);
            c_l2 += 0.5 * r[t] * r[t];
        }
        atomicAdd(cost, c_rob); atomicAdd(cost + 1, c_l2);
        if (j == i) continue;  // dqb_warp(dq_i, v_i) - dqb_warp(dq_i, v_i) == 0: no derivative
        add_block(H + (size_t)find_slot(row_ptr, col_idx, i, i) * 64, Ji, Ji, om);
        add_block(H + (size_t)find_slot(row_ptr, col_idx, i, j) * 64, Ji, Jj, om);
        add_block(H + (size_t)find_slot(row_ptr, col_idx, j, i) * 64, Jj, Ji, om);
        add_block(H + (size_t)find_slot(row_ptr, col_idx, j, j) * 64, Jj, Jj, om);
        for (int c = 0; c < 8; ++c) {
            atomicAdd(g + 8 * (size_t)i + c, om[0] * Ji[0][c] * r[0] + om[1] * Ji[1][c] * r[1] + om[2] * Ji[2][c] * r[2]);
            atomicAdd(g + 8 * (size_t)j + c, om[0] * Jj[0][c] * r[0] + om[1] * Jj[1][c] * r[1] + om[2] * Jj[2][c] * r[2]);
        }
    }
}

// global rigid dq: 8x8 normal equations, warp-shuffle + block reduction, one atomic per value per block
__global__ void __launch_bounds__(256) lw_normal_eq_kernel(const __grid_constant__ GNParams P, const double* dq, const __grid_constant__ LwArg L,
                                                          double* H8, double* g8, double* cost) {
    constexpr int NV = 36 + 8 + 2;
    double acc[NV];
    for (int t = 0; t < NV; ++t) acc[t] = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P.n_vert; i += (int64_t)gridDim.x * blockDim.x) {
        double r, J[8];
        lw_residual_jac(P, dq, L.lw, i, &r, J);
        const double om = huber_weight(r, P.huber, P.f_scale);
        int t = 0;
        for (int a = 0; a < 8; ++a)
            for (int b = a; b < 8; ++b) acc[t++] += om * J[a] * J[b];
        for (int a = 0; a < 8; ++a) acc[36 + a] += om * J[a] * r;
        acc[44] += huber_rho(r, P.huber, P.f_scale);
        acc[45] += 0.5 * r * r;
    }
    __shared__ double sm[8][NV];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int t = 0; t < NV; ++t) {
        double v = acc[t];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) sm[wid][t] = v;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double v = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += sm[w][threadIdx.x];
        const int t = threadIdx.x;
        if (t < 36) {
            int a = 0, rem = t;
            while (rem >= 8 - a) { rem -= 8 - a; ++a; }
            const int b = a + rem;
            atomicAdd(H8 + a * 8 + b, v);
            if (a != b) atomicAdd(H8 + b * 8 + a, v);
        } else if (t < 44) {
            atomicAdd(g8 + (t - 36), v);
        } else {
            atomicAdd(cost + (t - 44), v);
        }
    }
}

// ---- damped solve: (H + mu I) delta = -g, block-Jacobi preconditioned CG, all scalars stay on the device -----------
// S[0]=mu  S[1]=rz  S[2]=pq  S[3]=rz_new  S[4]=rz0  S[5]=done  S[6]=iterations  S[7]=trace
__global__ void trace_kernel(const int32_t* row_ptr, const int32_t* col_idx, const double* H, int n, double* S) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double t = 0.0;
    if (i < n) {
        const double* D = H + (size_t)find_slot(row_ptr, col_idx, i, i) * 64;
        for (int c = 0; c < 8; ++c) t += D[c * 9];
    }
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if ((threadIdx.x & 31) == 0 && t != 0.0) atomicAdd(S + 7, t);
}

__global__ void pcg_init_kernel(const int32_t* row_ptr, const int32_t* col_idx, const double* H, const double* g, int n, double lambda,
                                double* Minv, double* delta, double* r, double* z_out, double* S) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const double mu = lambda * S[7] / (8.0 * n);
    double part = 0.0;
    if (i < n) {
        const double* D = H + (size_t)find_slot(row_ptr, col_idx, i, i) * 64;
        double A[8][16];
        for (int a = 0; a < 8; ++a)
            for (int b = 0; b < 8; ++b) { A[a][b] = D[a * 8 + b] + (a == b ? mu : 0.0); A[a][8 + b] = (a == b) ? 1.0 : 0.0; }
        // Gauss-Jordan with partial pivoting (the block is SPD once damped; pivoting guards degenerate nodes)
        for (int c = 0; c < 8; ++c) {
            int piv = c;
            double best = fabs(A[c][c]);
            for (int a = c + 1; a < 8; ++a)
                if (fabs(A[a][c]) > best) { best = fabs(A[a][c]); piv = a; }
            if (piv != c)
                for (int b = 0; b < 16; ++b) { const double t = A[c][b]; A[c][b] = A[piv][b]; A[piv][b] = t; }
            const double d = A[c][c];
            const double inv = (d != 0.0) ? 1.0 / d : 0.0;
            for (int b = 0; b < 16; ++b) A[c][b] *= inv;
            for (int a = 0; a < 8; ++a) {
                if (a == c) continue;
                const double f = A[a][c];
                if (f != 0.0)
                    for (int b = 0; b < 16; ++b) A[a][b] -= f * A[c][b];
            }
        }
        double rr[8];
        for (int a = 0; a < 8; ++a) {
            rr[a] = -g[8 * (size_t)i + a];
            for (int b = 0; b < 8; ++b) Minv[(size_t)i * 64 + a * 8 + b] = A[a][8 + b];
        }
        for (int a = 0; a < 8; ++a) {
            double z = 0.0;
            for (int b = 0; b < 8; ++b) z += A[a][8 + b] * rr[b];
            delta[8 * (size_t)i + a] = 0.0;
            r[8 * (size_t)i + a] = rr[a];
            z_out[8 * (size_t)i + a] = z;
            part += rr[a] * z;
        }
    }
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((threadIdx.x & 31) == 0 && part != 0.0) { atomicAdd(S + 8, part); atomicAdd(S + 4, part); }
    if (i == 0) S[0] = mu;
}

// The whole PCG loop in ONE cooperative launch: the system is tiny (8N <= 32k rows, ~10 MB of L2-resident blocks), so
// separate launches are pure launch latency (4 dependent launches ~ 37 us/iteration measured); grid.sync() costs ~2 us.
// Chronopoulos-Gear arrangement of CG: with s = A z carried next to w = A p, both inner products of an iteration
// ((r,z) and (z,s)) are available after ONE matrix product, so an iteration is
//   phase 1 (node-local, 8 lanes per node): p = z + beta p, w = s + beta w, x += alpha p, r -= alpha w, z = Minv r, (r,z)
//   -- barrier --
//   phase 2 (one warp per block row):       s = (H + mu I) z, (z,s)
//   -- barrier --
// i.e. two grid-wide barriers per iteration instead of the three of the textbook ordering.  Every thread derives alpha and
// beta from the same globally reduced sums, so no scalar has to be broadcast.
__global__ void __launch_bounds__(256) pcg_global_kernel(const int32_t* row_ptr, const int32_t* col_idx, const double* H, const double* Minv, int n,
                                                        int max_iter, double tol2, double* delta, double* r, double* z, double* p, double* sv,
                                                        double* w, double* S) {
    // scalars: S[0]=mu  S[1]=final (r,z)  S[4]=(r,z) at start  S[6]=iterations;  G = S+8: (r,z), D = S+10: (z,s), both
    // double-buffered by iteration parity so that no "rotate the scalars" phase is needed.
    cg::grid_group grid = cg::this_grid();
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nthreads = gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    double* G = S + 8;
    double* D = S + 10;
    const double mu = S[0];
    const double rz0 = S[4];
    // s = (H + mu I) z and the partial (z,s) into *acc_zs.  One warp per block row: lane = (column group cgp = lane>>3,
    // row a = lane&7); the 8 lanes of a group read one 512-byte block row-by-row (coalesced), groups stride over the blocks.
    auto spmv = [&](double* acc_zs) {
        double part = 0.0;
        const int warp = tid >> 5, nwarps = nthreads >> 5;
        const int a = lane & 7, cgp = lane >> 3;
        for (int i = warp; i < n; i += nwarps) {
            double acc = 0.0;
            for (int t = row_ptr[i] + cgp; t < row_ptr[i + 1]; t += 4) {
                const double* Hb = H + (size_t)t * 64 + a * 8;
                const double* zj = z + 8 * (size_t)col_idx[t];
#pragma unroll
                for (int b = 0; b < 8; ++b) acc += Hb[b] * zj[b];
            }
            acc += __shfl_xor_sync(0xffffffffu, acc, 8);
            acc += __shfl_xor_sync(0xffffffffu, acc, 16);
            if (cgp == 0) {
                const double za = z[8 * (size_t)i + a];
                acc += mu * za;
                sv[8 * (size_t)i + a] = acc;
                part += acc * za;
            }
        }
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if (lane == 0 && part != 0.0) atomicAdd(acc_zs, part);
    };
    spmv(D + 0);
    grid.sync();
    double g_prev = 1.0, alpha_prev = 1.0;
    int it = 0;
    double g_last = G[0];
    for (; it < max_iter; ++it) {
        const int cur = it & 1, nxt = cur ^ 1;
        const double gam = G[cur], dzs = D[cur];
        const double beta = (it == 0 || g_prev == 0.0) ? 0.0 : gam / g_prev;
        const double den = (it == 0) ? dzs : dzs - beta * gam / alpha_prev;
        const double alpha = (den != 0.0) ? gam / den : 0.0;
        if (tid == 0) D[nxt] = 0.0;
        double part = 0.0;
        {
            const int grp = tid >> 3, ngrp = nthreads >> 3, a = lane & 7;
            const int nloop = (n + ngrp - 1) / ngrp;
            for (int l = 0; l < nloop; ++l) {
                const int i = grp + l * ngrp;
                const bool ok = i < n;
                const size_t t = 8 * (size_t)(ok ? i : 0) + a;
                double ra = 0.0;
                if (ok) {
                    const double pa = (it == 0) ? z[t] : z[t] + beta * p[t];
                    const double wa = (it == 0) ? sv[t] : sv[t] + beta * w[t];
                    p[t] = pa;
                    w[t] = wa;
                    delta[t] += alpha * pa;
                    ra = r[t] - alpha * wa;
                    r[t] = ra;
                }
                double zz = 0.0;
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    const double rb = __shfl_sync(0xffffffffu, ra, (lane & 24) | b);
                    if (ok) zz += Minv[(size_t)i * 64 + a * 8 + b] * rb;
                }
                if (ok) {
                    z[t] = zz;
                    part += ra * zz;
                }
            }
        }
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if (lane == 0 && part != 0.0) atomicAdd(G + nxt, part);
        grid.sync();
        g_last = G[nxt];
        if (!(g_last > tol2 * rz0)) { ++it; break; }   // identical on every thread
        if (tid == 0) G[cur] = 0.0;                   // accumulated again in phase 1 of the next iteration
        spmv(D + nxt);
        g_prev = gam;
        alpha_prev = alpha;
        grid.sync();
    }
    if (tid == 0) { S[6] = (double)it; S[1] = g_last; }
}

// Resident variant of the same iteration for systems that fit on the chip (N <= 4 rows per warp x 1184 warps): the matrix
// never changes during a solve, so every CTA copies the 8x8 blocks (and column indices, and inverse diagonal blocks) of the
// rows it owns into shared memory ONCE, and the node-local vectors r, p, w, delta, s, z of a row live in the registers of
// the warp that owns it for the whole solve.  Per iteration only z (64 bytes per node) goes through L2: written in phase 1,
// gathered by the neighbours' matrix products in phase 2.
template <int RPW>
__global__ void __launch_bounds__(256, 1) pcg_resident_kernel(const int32_t* row_ptr, const int32_t* col_idx, const double* H, const double* Minv,
                                                              int n, int max_iter, double tol2, double* delta, const double* r_in, double* z,
                                                              double* S, int cap_blocks) {
    extern __shared__ double smem[];
    cg::grid_group grid = cg::this_grid();
    const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
    const int a = lane & 7, cgp = lane >> 3;
    constexpr int ROWS_PER_CTA = 8 * RPW;
    double* Ms = smem;                                        // [ROWS_PER_CTA][64]
    // blocks are stored TRANSPOSED, columns swizzled by the block's parity: element (a, b) of cached block loc lives at
    // loc*64 + (b ^ (loc & 1))*8 + a.  The 8 lanes of a group then sweep 16 consecutive banks and the groups of a warp (which
    // read consecutive blocks) alternate between the two halves of the bank array -> the 2 wavefronts a 64-bit warp access
    // needs anyway (row-major blocks were a 16-way bank conflict: 2.2 -> 1.6 ms per solve).
    constexpr int HS = 64;
    double* Hs = Ms + ROWS_PER_CTA * 64;                      // [cap_blocks][HS]
    int32_t* cs = reinterpret_cast<int32_t*>(Hs + (size_t)cap_blocks * HS);   // [cap_blocks]
    const int row_first = blockIdx.x * ROWS_PER_CTA;
    const int row_end = min(row_first + ROWS_PER_CTA, n);
    const int bs = row_first < n ? row_ptr[row_first] : 0;
    const int be = row_first < n ? row_ptr[row_end] : 0;
    const int ncache = min(be - bs, cap_blocks);
    for (int t = threadIdx.x; t < ncache * 64; t += blockDim.x) Hs[(t >> 6) * HS + (((t & 7) ^ ((t >> 6) & 1)) * 8) + ((t >> 3) & 7)] = H[(size_t)bs * 64 + t];
    for (int t = threadIdx.x; t < ncache; t += blockDim.x) cs[t] = col_idx[bs + t];
    for (int t = threadIdx.x; t < (row_end - row_first) * 64; t += blockDim.x) Ms[t] = Minv[(size_t)row_first * 64 + t];
    __syncthreads();
    double* G = S + 8;
    double* D = S + 10;
    const double mu = S[0];
    const double rz0 = S[4];
    // registers of the lanes cgp == 0 (component a of each owned row)
    double rr[RPW], pp[RPW], ww[RPW], dd[RPW], ss[RPW], zz[RPW];
    int rs[RPW], re[RPW];
#pragma unroll
    for (int j = 0; j < RPW; ++j) {
        const int i = row_first + wic * RPW + j;
        const bool ok = i < n;
        rr[j] = ok ? r_in[8 * (size_t)i + a] : 0.0;
        zz[j] = ok ? z[8 * (size_t)i + a] : 0.0;
        pp[j] = 0.0; ww[j] = 0.0; dd[j] = 0.0; ss[j] = 0.0;
        rs[j] = ok ? row_ptr[i] : 0;
        re[j] = ok ? row_ptr[i + 1] : 0;
    }
    __shared__ double red[8];
    // one global atomic per CTA and inner product (the 8 warp sums meet in shared memory first)
    auto cta_add = [&](double part, double* target) {
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if (lane == 0) red[wic] = part;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
#pragma unroll
            for (int k = 0; k < 8; ++k) t += red[k];
            if (t != 0.0) atomicAdd(target, t);
        }
        __syncthreads();
    };
    // s = (H + mu I) z for the owned rows.  Lane (a, cgp) handles row a of the blocks cgp, cgp+4, ... of the block row.
    // Every lane gathers ONE component of the neighbour's z per block (the 8 lanes of a group cover the 8 components and
    // exchange them by shuffle), four blocks in flight, so a row of <= 16 blocks pays a single L2 round trip.
    auto gather = [&](int j, int t0, double* zv, const double** Hb, int* hst) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int t = t0 + cgp + 4 * u;
            const bool valid = t < re[j];
            const int loc = t - bs;
            const bool cached = valid && loc < ncache;
            const int col = valid ? (cached ? cs[loc] : col_idx[t]) : 0;
            zv[u] = valid ? z[8 * (size_t)col + a] : 0.0;
            Hb[u] = cached ? Hs + (size_t)loc * HS + a : valid ? H + (size_t)t * 64 + a * 8 : Hs;
            hst[u] = (cached || !valid) ? 8 | ((loc & 1) << 8) : 1;   // element (a, b) is Hb[(b ^ parity) * stride]; parity in bit 8
        }
    };
    auto consume = [&](const double* zv, const double* const* Hb, const int* hst, double acc) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int b = 0; b < 8; ++b)   // b runs over the PHYSICAL columns; the matching z component is b ^ parity
                acc += Hb[u][b * (hst[u] & 0xff)] * __shfl_sync(0xffffffffu, zv[u], (lane & 24) | (b ^ (hst[u] >> 8)));
        }
        return acc;
    };
    auto spmv = [&](double* acc_zs) {
        double part = 0.0;
        // the first 16 blocks of EVERY owned row are gathered before any arithmetic: one L2 round trip per product
        double zv0[RPW][4];
        const double* Hb0[RPW][4];
        int hs0[RPW][4];
#pragma unroll
        for (int j = 0; j < RPW; ++j) gather(j, rs[j], zv0[j], Hb0[j], hs0[j]);
#pragma unroll
        for (int j = 0; j < RPW; ++j) {
            double acc = consume(zv0[j], Hb0[j], hs0[j], 0.0);
            for (int t0 = rs[j] + 16; t0 < re[j]; t0 += 16) {     // rows longer than 16 blocks; uniform across the warp
                double zv[4];
                const double* Hb[4];
                int hst[4];
                gather(j, t0, zv, Hb, hst);
                acc = consume(zv, Hb, hst, acc);
            }
            acc += __shfl_xor_sync(0xffffffffu, acc, 8);
            acc += __shfl_xor_sync(0xffffffffu, acc, 16);
            acc += mu * zz[j];
            ss[j] = acc;
            if (cgp == 0 && rs[j] < re[j]) part += acc * zz[j];
        }
        cta_add(part, acc_zs);
    };
    spmv(D + 0);
    grid.sync();
    double g_prev = 1.0, alpha_prev = 1.0;
    int it = 0;
    double g_last = G[0];
    const bool leader = blockIdx.x == 0 && threadIdx.x == 0;
    for (; it < max_iter; ++it) {
        const int cur = it & 1, nxt = cur ^ 1;
        const double gam = G[cur], dzs = D[cur];
        const double beta = (it == 0 || g_prev == 0.0) ? 0.0 : gam / g_prev;
        const double den = (it == 0) ? dzs : dzs - beta * gam / alpha_prev;
        const double alpha = (den != 0.0) ? gam / den : 0.0;
        if (leader) D[nxt] = 0.0;
        double part = 0.0;
#pragma unroll
        for (int j = 0; j < RPW; ++j) {
            const int i = row_first + wic * RPW + j;
            pp[j] = (it == 0) ? zz[j] : zz[j] + beta * pp[j];
            ww[j] = (it == 0) ? ss[j] : ss[j] + beta * ww[j];
            dd[j] += alpha * pp[j];
            rr[j] -= alpha * ww[j];
            // z = Minv r: lane (a, cgp) multiplies columns 2cgp, 2cgp+1 of row a; r_b comes from lane b (a group-0 lane)
            const double* Mrow = Ms + (size_t)(wic * RPW + j) * 64 + a * 8 + 2 * cgp;
            const double r0 = __shfl_sync(0xffffffffu, rr[j], 2 * cgp), r1 = __shfl_sync(0xffffffffu, rr[j], 2 * cgp + 1);
            double zn = (i < n) ? Mrow[0] * r0 + Mrow[1] * r1 : 0.0;
            zn += __shfl_xor_sync(0xffffffffu, zn, 8);
            zn += __shfl_xor_sync(0xffffffffu, zn, 16);
            zz[j] = zn;
            if (cgp == 0 && i < n) {
                z[8 * (size_t)i + a] = zn;
                part += rr[j] * zn;
            }
        }
        cta_add(part, G + nxt);
        grid.sync();
        g_last = G[nxt];
        if (!(g_last > tol2 * rz0)) { ++it; break; }
        if (leader) G[cur] = 0.0;
        spmv(D + nxt);
        g_prev = gam;
        alpha_prev = alpha;
        grid.sync();
    }
#pragma unroll
    for (int j = 0; j < RPW; ++j) {
        const int i = row_first + wic * RPW + j;
        if (cgp == 0 && i < n) delta[8 * (size_t)i + a] = dd[j];
    }
    if (leader) { S[6] = (double)it; S[1] = g_last; }
}

// Pipelined arrangement (Ghysels & Vanroose) of the same preconditioned CG for the resident case: with u = Minv r, w = A u carried
// as recurrences next to m = Minv w, n = A m, both inner products of an iteration ((r,u) and (w,u)) are node-local sums over
// vectors that are already in registers, and the only thing the matrix product needs from the other CTAs is m.  An iteration is
//   local: (r,u), (w,u), m = Minv w -> L2          -- ONE barrier --          n = A m; z,q,s,p recurrences; x, r, u, w updates
// i.e. one grid-wide barrier per iteration (the barrier, not the arithmetic, is what an iteration of this small system costs:
// 6.2 us with two barriers).  m alternates between two buffers (a fast CTA writes m_{i+1} while a slow one still gathers m_i) and
// the sums rotate through three slots.  Same stopping rule, same iterates up to rounding (1e-7 relative on the bench system).
// Measured (B200, 993 nodes, 15331 blocks, 324 iterations): 1.56 ms against 2.00 ms for the two-barrier kernel.  Folding the two sums
// into a hand-made flag barrier ({data, epoch} words gathered by CTA 0 or by every CTA) was tried and is slower: 2.3 - 2.5 ms.
template <int RPW>
__global__ void __launch_bounds__(256, 1) pcg_pipelined_kernel(const int32_t* row_ptr, const int32_t* col_idx, const double* H, const double* Minv,
                                                               int n, int max_iter, double tol2, double* delta, const double* r_in, double* buf0,
                                                               double* buf1, double* S, int cap_blocks) {
    extern __shared__ double smem[];
    cg::grid_group grid = cg::this_grid();
    const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
    const int a = lane & 7, cgp = lane >> 3;
    constexpr int ROWS_PER_CTA = 8 * RPW;
    // blocks of a row gathered per round = 4 GU: every round is an L2 round trip on the critical path of the iteration, and rows of
    // a body graph with k = 4 hold up to ~20 blocks (mean 15): five per lane group settle a row in one round
    constexpr int GU = 5;
    double* Ms = smem;                                        // [ROWS_PER_CTA][64]
    constexpr int HS = 64;                                    // blocks transposed + swizzled as in pcg_resident_kernel
    double* Hs = Ms + ROWS_PER_CTA * 64;                      // [cap_blocks][HS]
    int32_t* cs = reinterpret_cast<int32_t*>(Hs + (size_t)cap_blocks * HS);   // [cap_blocks]
    const int row_first = blockIdx.x * ROWS_PER_CTA;
    const int row_end = min(row_first + ROWS_PER_CTA, n);
    const int bs = row_first < n ? row_ptr[row_first] : 0;
    const int be = row_first < n ? row_ptr[row_end] : 0;
    const int ncache = min(be - bs, cap_blocks);
    // One block row per warp and every block of the CTA cached (the usual case: <= 1184 nodes): the product runs without
    // shuffles.  Lane (a, bq) owns row a and the column pair (2bq, 2bq+1) of EVERY block of its row, reads the two matching
    // components of the neighbour's vector with one 16-byte load, and the four pair sums meet in two shuffles at the end.  Blocks
    // are stored as [column parity][bq][a] so that the 32 lanes of a 64-bit access sweep 32 consecutive doubles (2 wavefronts).
    const bool direct = RPW == 1 && be - bs <= cap_blocks;
    if (direct) {
        for (int t = threadIdx.x; t < ncache * 64; t += blockDim.x) {
            const int ra = (t >> 3) & 7, cb = t & 7;     // H is row-major: element (ra, cb) of block t >> 6
            Hs[(t >> 6) * HS + (cb & 1) * 32 + (cb >> 1) * 8 + ra] = H[(size_t)bs * 64 + t];
        }
    } else {
        for (int t = threadIdx.x; t < ncache * 64; t += blockDim.x) Hs[(t >> 6) * HS + (((t & 7) ^ ((t >> 6) & 1)) * 8) + ((t >> 3) & 7)] = H[(size_t)bs * 64 + t];
    }
    for (int t = threadIdx.x; t < ncache; t += blockDim.x) cs[t] = col_idx[bs + t];
    for (int t = threadIdx.x; t < (row_end - row_first) * 64; t += blockDim.x) Ms[t] = Minv[(size_t)row_first * 64 + t];
    __syncthreads();
    double* G = S + 8;     // (r,u) of iteration it in slot it % 3; slot 0 holds (r0,u0) from pcg_init_kernel
    double* D = S + 11;    // (w,u)
    const double mu = S[0];
    const double rz0 = S[4];
    double rr[RPW], uu[RPW], ww[RPW], mm[RPW], zz[RPW], qq[RPW], sv[RPW], pp[RPW], dd[RPW];
    int rs[RPW], re[RPW];
#pragma unroll
    for (int j = 0; j < RPW; ++j) {
        const int i = row_first + wic * RPW + j;
        const bool ok = i < n;
        rr[j] = ok ? r_in[8 * (size_t)i + a] : 0.0;
        uu[j] = ok ? buf0[8 * (size_t)i + a] : 0.0;
        ww[j] = 0.0; mm[j] = 0.0; zz[j] = 0.0; qq[j] = 0.0; sv[j] = 0.0; pp[j] = 0.0; dd[j] = 0.0;
        rs[j] = ok ? row_ptr[i] : 0;
        re[j] = ok ? row_ptr[i + 1] : 0;
    }
    __shared__ double red[16];
    auto cta_add2 = [&](double p0, double p1, double* t0, double* t1) {
        for (int o = 16; o > 0; o >>= 1) { p0 += __shfl_xor_sync(0xffffffffu, p0, o); p1 += __shfl_xor_sync(0xffffffffu, p1, o); }
        if (lane == 0) { red[wic] = p0; red[8 + wic] = p1; }
        __syncthreads();
        if (threadIdx.x < 2) {
            double t = 0.0;
#pragma unroll
            for (int k = 0; k < 8; ++k) t += red[8 * threadIdx.x + k];
            if (t != 0.0) atomicAdd(threadIdx.x ? t1 : t0, t);
        }
        __syncthreads();
    };
    auto gather = [&](const double* src, int j, int t0, double* zv, const double** Hb, int* hst) {
#pragma unroll
        for (int u = 0; u < GU; ++u) {
            const int t = t0 + cgp + 4 * u;
            const bool valid = t < re[j];
            const int loc = t - bs;
            const bool cached = valid && loc < ncache;
            const int col = valid ? (cached ? cs[loc] : col_idx[t]) : 0;
            zv[u] = valid ? __ldcg(src + 8 * (size_t)col + a) : 0.0;
            Hb[u] = cached ? Hs + (size_t)loc * HS + a : valid ? H + (size_t)t * 64 + a * 8 : Hs;
            hst[u] = (cached || !valid) ? 8 | ((loc & 1) << 8) : 1;
        }
    };
    auto consume = [&](const double* zv, const double* const* Hb, const int* hst, double acc) {
        double acc1 = 0.0;   // two chains: the sum is a string of dependent fp64 FMAs otherwise
#pragma unroll
        for (int u = 0; u < GU; ++u) {
#pragma unroll
            for (int b = 0; b < 8; b += 2) {
                acc += Hb[u][b * (hst[u] & 0xff)] * __shfl_sync(0xffffffffu, zv[u], (lane & 24) | (b ^ (hst[u] >> 8)));
                acc1 += Hb[u][(b + 1) * (hst[u] & 0xff)] * __shfl_sync(0xffffffffu, zv[u], (lane & 24) | ((b + 1) ^ (hst[u] >> 8)));
            }
        }
        return acc + acc1;
    };
    // out = (H + mu I) v for the owned rows; v of the other rows is gathered from `src`, the own rows' v is `own`
    auto spmv = [&](const double* src, const double* own, double* out) {
        if (RPW == 1 && direct) {
            constexpr int GD = 20;                       // blocks in flight per round
            const int bq = cgp;
            double acc0 = 0.0, acc1 = 0.0;
            for (int t0 = rs[0]; t0 < re[0]; t0 += GD) {
                double2 zr[GD];
#pragma unroll
                for (int u = 0; u < GD; ++u) {
                    const int t = t0 + u;
                    zr[u] = t < re[0] ? __ldcg(reinterpret_cast<const double2*>(src + 8 * (size_t)cs[t - bs]) + bq) : make_double2(0.0, 0.0);
                }
#pragma unroll
                for (int u = 0; u < GD; ++u) {
                    const int t = t0 + u;
                    if (t < re[0]) {
                        const double* hb = Hs + (size_t)(t - bs) * HS + bq * 8 + a;
                        acc0 += hb[0] * zr[u].x;
                        acc1 += hb[32] * zr[u].y;
                    }
                }
            }
            double acc = acc0 + acc1;
            acc += __shfl_xor_sync(0xffffffffu, acc, 8);
            acc += __shfl_xor_sync(0xffffffffu, acc, 16);
            out[0] = acc + mu * own[0];
            return;
        }
        double zv0[RPW][GU];
        const double* Hb0[RPW][GU];
        int hs0[RPW][GU];
#pragma unroll
        for (int j = 0; j < RPW; ++j) gather(src, j, rs[j], zv0[j], Hb0[j], hs0[j]);
#pragma unroll
        for (int j = 0; j < RPW; ++j) {
            double acc = consume(zv0[j], Hb0[j], hs0[j], 0.0);
            for (int t0 = rs[j] + 4 * GU; t0 < re[j]; t0 += 4 * GU) {   // rows longer than one round; uniform across the warp
                double zv[GU];
                const double* Hb[GU];
                int hst[GU];
                gather(src, j, t0, zv, Hb, hst);
                acc = consume(zv, Hb, hst, acc);
            }
            acc += __shfl_xor_sync(0xffffffffu, acc, 8);
            acc += __shfl_xor_sync(0xffffffffu, acc, 16);
            out[j] = acc + mu * own[j];
        }
    };
    spmv(buf0, uu, ww);   // w0 = A u0 (u0 written by pcg_init_kernel, an earlier launch)
    double g_prev = 1.0, alpha_prev = 1.0;
    int it = 0;
    double g_last = G[0];
    const bool leader = blockIdx.x == 0 && threadIdx.x == 0;
    for (; it < max_iter; ++it) {
        const int slot = it % 3, nslot = (it + 1) % 3;
        double* mdst = (it & 1) ? buf0 : buf1;
        double part_g = 0.0, part_d = 0.0;
#pragma unroll
        for (int j = 0; j < RPW; ++j) {
            const int i = row_first + wic * RPW + j;
            // m = Minv w: lane (a, cgp) multiplies columns 2cgp, 2cgp+1 of row a
            const double* Mrow = Ms + (size_t)(wic * RPW + j) * 64 + a * 8 + 2 * cgp;
            const double w0 = __shfl_sync(0xffffffffu, ww[j], 2 * cgp), w1 = __shfl_sync(0xffffffffu, ww[j], 2 * cgp + 1);
            double mn = (i < n) ? Mrow[0] * w0 + Mrow[1] * w1 : 0.0;
            mn += __shfl_xor_sync(0xffffffffu, mn, 8);
            mn += __shfl_xor_sync(0xffffffffu, mn, 16);
            mm[j] = mn;
            if (cgp == 0 && i < n) {
                __stcg(mdst + 8 * (size_t)i + a, mn);
                part_g += rr[j] * uu[j];
                part_d += ww[j] * uu[j];
            }
        }
        if (leader) { G[nslot] = 0.0; D[nslot] = 0.0; }
        cta_add2(it == 0 ? 0.0 : part_g, part_d, G + slot, D + slot);
        grid.sync();
        const double gam = __ldcg(G + slot), dl = __ldcg(D + slot);
        g_last = gam;
        if (!(gam > tol2 * rz0)) break;
        double nn[RPW];
        spmv(mdst, mm, nn);
        const double beta = (it == 0 || g_prev == 0.0) ? 0.0 : gam / g_prev;
        const double den = (it == 0) ? dl : dl - beta * gam / alpha_prev;
        const double alpha = (den != 0.0) ? gam / den : 0.0;
#pragma unroll
        for (int j = 0; j < RPW; ++j) {
            zz[j] = nn[j] + beta * zz[j];
            qq[j] = mm[j] + beta * qq[j];
            sv[j] = ww[j] + beta * sv[j];
            pp[j] = uu[j] + beta * pp[j];
            dd[j] += alpha * pp[j];
            rr[j] -= alpha * sv[j];
            uu[j] -= alpha * qq[j];
            ww[j] -= alpha * zz[j];
        }
        g_prev = gam;
        alpha_prev = alpha;
    }
#pragma unroll
    for (int j = 0; j < RPW; ++j) {
        const int i = row_first + wic * RPW + j;
        if (cgp == 0 && i < n) delta[8 * (size_t)i + a] = dd[j];
    }
    if (leader) { S[6] = (double)it; S[1] = g_last; }
}

__global__ void apply_delta_kernel(const double* x, const double* delta, int n8, double* x_new) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n8) x_new[t] = x[t] + delta[t];
}

int blocks_for(int64_t n, int threads, int cap = 148 * 32) {
    int64_t b = (n + threads - 1) / threads;
    if (b < 1) b = 1;
    return (int)(b > cap ? cap : b);
}

}  // namespace

// ===================================================================================================================
extern "C" int dfb_gn_residuals(const dfb_gn_problem* prob, const double* x, int x_is_f32, double* f_out, dfb_stream_t stream) {
    if (int r = validate(prob)) return r;
    DFB_REQUIRE(x && f_out, "null pointer");
    const GNParams P = to_params(prob);
    const int64_t n = P.n_vert + (int64_t)P.n_nodes * P.k;
    residual_kernel<<<blocks_for(n, 128), 128, 0, (cudaStream_t)stream>>>(P, x, x_is_f32, f_out);
    DFB_LAUNCH_CHECK("residual_kernel");
    return DFB_OK;
}

extern "C" int dfb_gn_residuals_lw(const dfb_gn_problem* prob, const double* node_dq, int dq_is_f32, const double* lw, int lw_is_f32,
                                   double* f_out, dfb_stream_t stream) {
    if (int r = validate(prob)) return r;
    DFB_REQUIRE(node_dq && lw && f_out, "null pointer");
    const GNParams P = to_params(prob);
    LwArg L;
    for (int i = 0; i < 8; ++i) L.lw[i] = lw[i];
    L.is_f32 = lw_is_f32;
    if (P.n_vert == 0) return DFB_OK;
    residual_lw_kernel<<<blocks_for(P.n_vert, 128), 128, 0, (cudaStream_t)stream>>>(P, node_dq, dq_is_f32, L, f_out);
    DFB_LAUNCH_CHECK("residual_lw_kernel");
    return DFB_OK;
}

extern "C" int dfb_gn_pattern_rows(const dfb_gn_problem* prob, uint32_t* bitmap, int32_t* row_ptr, dfb_stream_t stream) {
    if (int r = validate(prob)) return r;
    DFB_REQUIRE(bitmap && row_ptr, "null pointer");
    const GNParams P = to_params(prob);
    cudaStream_t s = (cudaStream_t)stream;
    const int words = (P.n_nodes + 31) / 32;
    DFB_CUDA(cudaMemsetAsync(bitmap, 0, (size_t)P.n_nodes * words * sizeof(uint32_t), s));
    const int64_t n = P.n_vert + (int64_t)P.n_nodes * (P.k + 1);
    pattern_mark_kernel<<<blocks_for(n, 256), 256, 0, s>>>(P, bitmap, words);
    DFB_LAUNCH_CHECK("pattern_mark_kernel");
    pattern_count_kernel<<<(P.n_nodes + 127) / 128, 128, 0, s>>>(bitmap, P.n_nodes, words, row_ptr);
    DFB_LAUNCH_CHECK("pattern_count_kernel");
    pattern_scan_kernel<<<1, 1024, 0, s>>>(row_ptr, P.n_nodes);
    DFB_LAUNCH_CHECK("pattern_scan_kernel");
    return DFB_OK;
}

extern "C" int dfb_gn_pattern_cols(int n_nodes, const uint32_t* bitmap, const int32_t* row_ptr, int32_t* col_idx, dfb_stream_t stream) {
    DFB_REQUIRE(bitmap && row_ptr && col_idx && n_nodes > 0, "bad arguments");
    pattern_fill_kernel<<<(n_nodes + 127) / 128, 128, 0, (cudaStream_t)stream>>>(bitmap, n_nodes, (n_nodes + 31) / 32, row_ptr, col_idx);
    DFB_LAUNCH_CHECK("pattern_fill_kernel");
    return DFB_OK;
}

extern "C" int dfb_gn_normal_eq(const dfb_gn_problem* prob, const double* x, const int32_t* row_ptr, const int32_t* col_idx, int64_t nnzb,
                                double* H, double* g, double* cost, dfb_stream_t stream) {
    if (int r = validate(prob)) return r;
    DFB_REQUIRE(x && row_ptr && col_idx && H && g && cost && nnzb > 0, "bad arguments");
    const GNParams P = to_params(prob);
    cudaStream_t s = (cudaStream_t)stream;
    DFB_CUDA(cudaMemsetAsync(H, 0, (size_t)nnzb * 64 * sizeof(double), s));
    DFB_CUDA(cudaMemsetAsync(g, 0, (size_t)P.n_nodes * 8 * sizeof(double), s));
    DFB_CUDA(cudaMemsetAsync(cost, 0, 2 * sizeof(double), s));
    if (P.n_vert > 0) {
        const int nblk = blocks_for(P.n_vert, 256, 148 * 8);
        switch (P.k) {
#define DFB_NEQ_CASE(K_) case K_: normal_eq_data_kernel<K_><<<nblk, 256, 0, s>>>(P, x, row_ptr, col_idx, H, g, cost); break;
            DFB_NEQ_CASE(1) DFB_NEQ_CASE(2) DFB_NEQ_CASE(3) DFB_NEQ_CASE(4) DFB_NEQ_CASE(5) DFB_NEQ_CASE(6) DFB_NEQ_CASE(7) DFB_NEQ_CASE(8)
#undef DFB_NEQ_CASE
        }
        DFB_LAUNCH_CHECK("normal_eq_data_kernel");
        mirror_lower_kernel<<<(unsigned)((nnzb * 64 + 255) / 256), 256, 0, s>>>(row_ptr, col_idx, P.n_nodes, H);
        DFB_LAUNCH_CHECK("mirror_lower_kernel");
    }
    normal_eq_reg_kernel<<<blocks_for((int64_t)P.n_nodes * P.k, 64), 64, 0, s>>>(P, x, row_ptr, col_idx, H, g, cost);
    DFB_LAUNCH_CHECK("normal_eq_reg_kernel");
    return DFB_OK;
}

extern "C" int dfb_gn_lw_normal_eq(const dfb_gn_problem* prob, const double* node_dq, const double* lw, double* H8, double* g8, double* cost,
                                   dfb_stream_t stream) {
    if (int r = validate(prob)) return r;
    DFB_REQUIRE(node_dq && lw && H8 && g8 && cost, "null pointer");
    const GNParams P = to_params(prob);
    LwArg L;
    for (int i = 0; i < 8; ++i) L.lw[i] = lw[i];
    L.is_f32 = 0;
    cudaStream_t s = (cudaStream_t)stream;
    DFB_CUDA(cudaMemsetAsync(H8, 0, 64 * sizeof(double), s));
    DFB_CUDA(cudaMemsetAsync(g8, 0, 8 * sizeof(double), s));
    DFB_CUDA(cudaMemsetAsync(cost, 0, 2 * sizeof(double), s));
    if (P.n_vert == 0) return DFB_OK;
    lw_normal_eq_kernel<<<blocks_for(P.n_vert, 256, 148 * 4), 256, 0, s>>>(P, node_dq, L, H8, g8, cost);
    DFB_LAUNCH_CHECK("lw_normal_eq_kernel");
    return DFB_OK;
}

extern "C" int64_t dfb_gn_solve_workspace_doubles(int n_nodes) { return (int64_t)n_nodes * (64 + 8 * 5) + 16; }

extern "C" int dfb_gn_solve(int n_nodes, const int32_t* row_ptr, const int32_t* col_idx, const double* H, const double* g, double lambda,
                            int max_iter, double tol, const double* x, double* x_new, double* delta, double* workspace,
                            dfb_stream_t stream) {
    DFB_REQUIRE(n_nodes > 0 && row_ptr && col_idx && H && g && x && x_new && delta && workspace, "bad arguments");
    DFB_REQUIRE(max_iter >= 1 && lambda >= 0 && tol >= 0, "bad solver parameters");
    cudaStream_t s = (cudaStream_t)stream;
    const int n = n_nodes;
    double* S = workspace;
    double* Minv = workspace + 16;
    double* r = Minv + (size_t)n * 64;
    double* z = r + (size_t)n * 8;
    double* p = z + (size_t)n * 8;
    double* sv = p + (size_t)n * 8;
    double* w = sv + (size_t)n * 8;
    DFB_CUDA(cudaMemsetAsync(S, 0, 16 * sizeof(double), s));
    const int nb_node = (n + 127) / 128, nb_row = (8 * n + 127) / 128;
    trace_kernel<<<nb_node, 128, 0, s>>>(row_ptr, col_idx, H, n, S);
    pcg_init_kernel<<<nb_node, 128, 0, s>>>(row_ptr, col_idx, H, g, n, lambda, Minv, delta, r, z, S);
    DFB_LAUNCH_CHECK("pcg_init_kernel");
    const double tol2 = tol * tol;
    {
        static int sms = -1, smem_max = 0;
        if (sms < 0) {
            int dev = 0;
            DFB_CUDA(cudaGetDevice(&dev));
            DFB_CUDA(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
            DFB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        }
        static const bool force_global = getenv("DFB_PCG_GLOBAL") != nullptr;
        const int warps = sms * 8;
        static const int rpw_min = getenv("DFB_PCG_RPW") ? atoi(getenv("DFB_PCG_RPW")) : 1;   // tuning: fewer, fatter CTAs at the barrier
        const int rpw = max((n + warps - 1) / warps, min(rpw_min, 4));
        if (rpw <= 4 && !force_global) {
            // resident path: one CTA of 8 warps per SM, rpw block rows per warp, matrix blocks cached in shared memory
            const int rows_per_cta = 8 * rpw;
            const int blocks = (n + rows_per_cta - 1) / rows_per_cta;
            const size_t fixed = (size_t)rows_per_cta * 64 * sizeof(double);
            int cap_blocks = (int)(((size_t)smem_max - 1024 - fixed) / (64 * sizeof(double) + sizeof(int32_t)));
            const size_t smem = fixed + (size_t)cap_blocks * (64 * sizeof(double) + sizeof(int32_t));
            const double* r_in = r;
            static const bool two_barriers = getenv("DFB_PCG_TWO_BARRIERS") != nullptr;   // A/B: the Chronopoulos-Gear kernel
            if (two_barriers) {
                void* args[] = {(void*)&row_ptr, (void*)&col_idx, (void*)&H, (void*)&Minv, (void*)&n, (void*)&max_iter, (void*)&tol2,
                                (void*)&delta, (void*)&r_in, (void*)&z, (void*)&S, (void*)&cap_blocks};
                void* fn = rpw == 1 ? (void*)pcg_resident_kernel<1> : rpw == 2 ? (void*)pcg_resident_kernel<2>
                         : rpw == 3 ? (void*)pcg_resident_kernel<3> : (void*)pcg_resident_kernel<4>;
                DFB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                DFB_CUDA(cudaLaunchCooperativeKernel(fn, dim3(blocks), dim3(256), args, smem, s));
            } else {
                void* args[] = {(void*)&row_ptr, (void*)&col_idx, (void*)&H, (void*)&Minv, (void*)&n, (void*)&max_iter, (void*)&tol2,
                                (void*)&delta, (void*)&r_in, (void*)&z, (void*)&p, (void*)&S, (void*)&cap_blocks};
                void* fn = rpw == 1 ? (void*)pcg_pipelined_kernel<1> : rpw == 2 ? (void*)pcg_pipelined_kernel<2>
                         : rpw == 3 ? (void*)pcg_pipelined_kernel<3> : (void*)pcg_pipelined_kernel<4>;
                DFB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                DFB_CUDA(cudaLaunchCooperativeKernel(fn, dim3(blocks), dim3(256), args, smem, s));
            }
        } else {
            static int coop_blocks = -1;   // co-resident CTAs of pcg_global_kernel on this device
            if (coop_blocks < 0) {
                int per_sm = 0;
                DFB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pcg_global_kernel, 256, 0));
                coop_blocks = sms * (per_sm > 0 ? 1 : 0);
            }
            DFB_REQUIRE(coop_blocks > 0, "cooperative launch not possible on this device");
            int blocks = (32 * n + 255) / 256;   // one warp per block row
            if (blocks > coop_blocks) blocks = coop_blocks;
            void* args[] = {(void*)&row_ptr, (void*)&col_idx, (void*)&H, (void*)&Minv, (void*)&n, (void*)&max_iter, (void*)&tol2,
                            (void*)&delta, (void*)&r, (void*)&z, (void*)&p, (void*)&sv, (void*)&w, (void*)&S};
            DFB_CUDA(cudaLaunchCooperativeKernel((void*)pcg_global_kernel, dim3(blocks), dim3(256), args, 0, s));
        }
    }
    apply_delta_kernel<<<nb_row, 128, 0, s>>>(x, delta, 8 * n, x_new);
    DFB_LAUNCH_CHECK("apply_delta_kernel");
    return DFB_OK;
}
