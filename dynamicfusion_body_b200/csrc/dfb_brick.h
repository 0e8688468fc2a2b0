// dfb_brick.h -- brick-level conservative classification for the projective TSDF passes (a2/a3).
//
// A brick is 4 x 4 x 32 voxels (x,y,z): sixteen 128-byte rows of the z-fastest volume.  Per frame every brick is
// classified as
//     SKIP   : no voxel of the brick can be updated by any view          -> no memory traffic at all
//     CLAMP  : every voxel is certainly updated with min(tdist, tl) = tdist by a fixed set of views
//                                                                         -> pure streaming pass (16 B/voxel)
//     MIXED  : anything else                                              -> per-voxel kernel (dfb_voxel.h)
//
// Rigour.  For a voxel x with neighbours S and ANY positive blend weights w_i, the reference's warped point is
//     p'(x) = Q(b,x)/|b|^2,  b = sum_i w_i q_i      (Q2: 8-norm normalisation; Q = dqb_warp's quadratic form)
//           = sum_ij (w_i w_j <q_i,q_j>) P_ij(x) / sum_ij (w_i w_j <q_i,q_j>),   P_ij(x) = Qpol(q_i,q_j,x)/<q_i,q_j>
// (only pairs i,j that occur together in the kNN set of at least one voxel of the brick matter; they are cached as a
// bit mask over the lower triangle of the candidate list)
// where Qpol is the polarisation of Q.  If <q_i,q_j> > 0 for all pairs, p'(x) is a CONVEX COMBINATION of the points
// P_ij(x), and each P_ij is an affine map of x.  Hence p'(brick) lies in the bounding box of the boxes
// P_ij(brick), i,j in C, where C is the union of the kNN sets of the brick's voxels (cached per graph revision).
// The bound needs no knowledge of the weights -- it holds for the reference's float32/powf/exp-evaluated ones as well.
// The box is pushed through lw / extrinsics / intrinsics with interval arithmetic, the depth image is scanned over
// the resulting pixel rectangle, and the decision is taken only if it holds for the whole box (plus rounding slack).
#pragma once
#include "dfb_voxel.h"

namespace dfb {

constexpr int BRICK_X = 4, BRICK_Y = 4, BRICK_Z = 32;
constexpr int BRICK_MAXC = 24;        // cached candidate nodes per brick (more -> always MIXED)
constexpr int REGION_X = 16, REGION_Y = 16, REGION_Z = 32;   // 4 x 4 x 1 bricks
constexpr int REGION_MAXC = 64;                              // distinct nodes cached per region (more -> region unusable)
constexpr int REGION_PAIR_WORDS = 65;                        // 64*65/2 = 2080 pair bits
constexpr int REGION_REC_FLOATS = 32;                        // P_ref (row-major 3x4), D[3], code | view 0's QuadView: M (3x4), d[3], unused
constexpr int BRICK_PAIR_WORDS = 10;   // bit p = i*(i+1)/2 + j (j <= i) of the 24*25/2 = 300 candidate pairs
constexpr int BRICK_CLS_MIXED = 0xFF;
constexpr int REGION_MAX_RECT = 4096;   // depth pixels scanned for a whole region (one warp)
constexpr int BRICK_MAX_RECT = 512;   // depth pixels scanned per brick and view before giving up

struct Box3 { float lo[3], hi[3]; };

// Class of a brick (or region).  Views are classified one by one: `clamp` = views in which every voxel gets the clamped
// update, `frus` = views whose image certainly contains every voxel, `mixed` = views the box test could not settle (their
// bits in clamp / frus are 0).  cls = BRICK_CLS_MIXED when any view is unsettled, else the clamp mask (0 = SKIP).  The
// per-voxel tier of a MIXED brick only evaluates the unsettled views and takes the others from (clamp, frus).
struct BrickClass {
    int cls, clamp, frus, mixed;
};
DFB_HD BrickClass brick_class_all_mixed(int n_views) {
    BrickClass r;
    r.cls = BRICK_CLS_MIXED; r.clamp = 0; r.frus = 0; r.mixed = (1 << n_views) - 1;
    return r;
}

// affine map (row-major 3x4) of dqb_warp(q, .) for a possibly non-unit q, fp32
DFB_HD void dq_affine_f(const float* q, float* A) {
    const float w = q[0], x = q[1], y = q[2], z = q[3], dw = q[4], dx = q[5], dy = q[6], dz = q[7];
    const float s = w * w - (x * x + y * y + z * z);
    A[0] = s + 2 * x * x;        A[1] = 2 * (x * y - w * z);  A[2] = 2 * (x * z + w * y);
    A[4] = 2 * (x * y + w * z);  A[5] = s + 2 * y * y;        A[6] = 2 * (y * z - w * x);
    A[8] = 2 * (x * z - w * y);  A[9] = 2 * (y * z + w * x);  A[10] = s + 2 * z * z;
    A[3] = 2 * (w * dx - dw * x + (y * dz - z * dy));
    A[7] = 2 * (w * dy - dw * y + (z * dx - x * dz));
    A[11] = 2 * (w * dz - dw * z + (x * dy - y * dx));
}

// 2 * (polarisation of dqb_warp's quadratic form): the affine map x -> 2*Qpol(a, b, x); for a == b it is 2*A(a)
DFB_HD void dq_affine_polar2(const float* a, const float* b, float* A) {
    const float wa = a[0], xa = a[1], ya = a[2], za = a[3], dwa = a[4], dxa = a[5], dya = a[6], dza = a[7];
    const float wb = b[0], xb = b[1], yb = b[2], zb = b[3], dwb = b[4], dxb = b[5], dyb = b[6], dzb = b[7];
    const float s = 2.f * (wa * wb - (xa * xb + ya * yb + za * zb));
    const float xy = xa * yb + ya * xb, xz = xa * zb + za * xb, yz = ya * zb + za * yb;
    const float wx = wa * xb + xa * wb, wy = wa * yb + ya * wb, wz = wa * zb + za * wb;
    A[0] = s + 4.f * xa * xb;   A[1] = 2.f * (xy - wz);       A[2] = 2.f * (xz + wy);
    A[4] = 2.f * (xy + wz);     A[5] = s + 4.f * ya * yb;     A[6] = 2.f * (yz - wx);
    A[8] = 2.f * (xz - wy);     A[9] = 2.f * (yz + wx);       A[10] = s + 4.f * za * zb;
    A[3] = 2.f * ((wa * dxb + wb * dxa) - (dwa * xb + dwb * xa) + (ya * dzb + yb * dza) - (za * dyb + zb * dya));
    A[7] = 2.f * ((wa * dyb + wb * dya) - (dwa * yb + dwb * ya) + (za * dxb + zb * dxa) - (xa * dzb + xb * dza));
    A[11] = 2.f * ((wa * dzb + wb * dza) - (dwa * zb + dwb * za) + (xa * dyb + xb * dya) - (ya * dxb + yb * dxa));
}

DFB_HD void box_extend_affine(const float* A, float inv, const float* c, const float* h, Box3& b) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const float m = (A[4 * r] * c[0] + A[4 * r + 1] * c[1] + A[4 * r + 2] * c[2] + A[4 * r + 3]) * inv;
        const float rad = (fabsf(A[4 * r]) * h[0] + fabsf(A[4 * r + 1]) * h[1] + fabsf(A[4 * r + 2]) * h[2]) * fabsf(inv);
        const float slack = 4e-6f * (fabsf(m) + rad) + 1e-6f;
        b.lo[r] = fminf(b.lo[r], m - rad - slack);
        b.hi[r] = fmaxf(b.hi[r], m + rad + slack);
    }
}

// interval of row . [p,1] over the box (centre c, half h), with rounding slack
DFB_HD void row_interval(const float* row, const float* c, const float* h, float& lo, float& hi) {
    const float m = row[0] * c[0] + row[1] * c[1] + row[2] * c[2] + row[3];
    const float rad = fabsf(row[0]) * h[0] + fabsf(row[1]) * h[1] + fabsf(row[2]) * h[2];
    const float mag = fabsf(row[0] * c[0]) + fabsf(row[1] * c[1]) + fabsf(row[2] * c[2]) + fabsf(row[3]);
    const float slack = 2e-6f * (mag + rad) + 1e-6f;
    lo = m - rad - slack;
    hi = m + rad + slack;
}

DFB_HD void div_interval(float nlo, float nhi, float dlo, float dhi, float& lo, float& hi) {  // requires dlo > 0
    const float a = nlo / dlo, b = nlo / dhi, c = nhi / dlo, d = nhi / dhi;
    lo = fminf(fminf(a, b), fminf(c, d));
    hi = fmaxf(fmaxf(a, b), fmaxf(c, d));
    const float slack = 4e-6f * fmaxf(fabsf(lo), fabsf(hi)) + 1e-5f;
    lo -= slack;
    hi += slack;
}

// position of the n-th (0-based) set bit of w
DFB_HD int nth_set_bit(uint32_t w, int n) {
#if defined(__CUDA_ARCH__)
    return (int)__fns(w, 0, n + 1);
#else
    for (int b = 0; b < 32; ++b)
        if ((w >> b) & 1u) { if (n == 0) return b; --n; }
    return 32;
#endif
}

DFB_HD int popc32(uint32_t w) {
#if defined(__CUDA_ARCH__)
    return __popc(w);
#else
    return __builtin_popcount(w);
#endif
}

// ---- regions: one reference affine map + a rigorous deviation bound per 16x16x32 block of voxels ---------------------
// For every voxel x of region R the warped point is a convex combination of P_ij(x) over the node pairs (i,j) that share
// x's kNN set, hence  |p'(x) - P_ref(x)|_inf <= D_R := max over the region's co-occurring pairs of sup_{x in R} |P_ij(x) - P_ref(x)|_inf
// for ANY fixed affine P_ref (the diagonal map of the region's first node is used).  P_ij - P_ref is affine, so the sup over
// the region box is |dA c + dt| + |dA| h per axis: exact interval arithmetic, no dependency problem.  With (P_ref, D_R) the
// brick classifier needs one affine map per brick instead of a loop over its candidate pairs.  (D_R is the genuine spread
// of the pair maps -- up to a few voxels where distant nodes blend -- so it is only used to sort bricks into SKIP / CLAMP /
// MIXED; voxels of MIXED bricks are still classified by the pointwise DQB tier.)
// Reference map of a region: the MEAN of its nodes' own maps A(q_i) / |q_i|^2 (summed in list order).  Any fixed affine map is a
// valid reference; against taking one node's map the mean shrinks the deviation bound by a fifth (median D of the MIXED regions of
// the 256^3 benchmark scene 1.56 -> 1.25 voxels).  node_affine_normalised: one node's contribution; false when its dq is (nearly) zero.
DFB_HD bool node_affine_normalised(const float* qi, float* A) {
    float n0 = 0.f;
    for (int t = 0; t < 8; ++t) n0 += qi[t] * qi[t];
    dq_affine_f(qi, A);
    const bool ok = n0 > 1e-20f;
    const float inv = ok ? 1.0f / n0 : 0.f;
    for (int t = 0; t < 12; ++t) A[t] *= inv;
    return ok;
}
DFB_HDN bool region_reference_map(const float* q, int cnt, float* Pref) {
    for (int t = 0; t < 12; ++t) Pref[t] = 0.f;
    bool ok = true;
    for (int n = 0; n < cnt; ++n) {
        float A[12];
        if (!node_affine_normalised(q + 8 * n, A)) ok = false;
        for (int t = 0; t < 12; ++t) Pref[t] += A[t];
    }
    const float invc = 1.0f / (float)(cnt > 0 ? cnt : 1);
    for (int t = 0; t < 12; ++t) Pref[t] *= invc;
    return ok;
}

// region_pair_bound: contribution of one pair (local indices i >= j into `q`, the region's node dq list) to D_R.
DFB_HD bool region_pair_bound(const float* qi, const float* qj, bool diag, const float* Pref, const float* c, const float* h, float* dev) {
    float ni = 0.f, nj = 0.f, ip = 0.f;
    for (int t = 0; t < 8; ++t) { ni += qi[t] * qi[t]; nj += qj[t] * qj[t]; ip += qi[t] * qj[t]; }
    if (!(ni > 1e-20f) || !(nj > 1e-20f)) return false;
    if (!diag && !(ip > 0.25f * sqrtf(ni * nj))) return false;
    float Ap[12];
    dq_affine_polar2(qi, qj, Ap);
    const float inv = 0.5f / ip;
    for (int r = 0; r < 3; ++r) {
        float mid = 0.f, rad = 0.f, mag = 0.f;
        for (int t = 0; t < 3; ++t) {
            const float d = Ap[4 * r + t] * inv - Pref[4 * r + t];
            mid += d * c[t];
            rad += fabsf(d) * h[t];
            mag += fabsf(Ap[4 * r + t] * inv * c[t]) + fabsf(Pref[4 * r + t] * c[t]);
        }
        const float dt = Ap[4 * r + 3] * inv - Pref[4 * r + 3];
        mag += fabsf(Ap[4 * r + 3] * inv) + fabsf(Pref[4 * r + 3]);
        dev[r] = fmaxf(dev[r], fabsf(mid + dt) + rad + 4e-6f * (mag + rad) + 1e-6f);
    }
    return true;
}

// Execution context: the same code runs with one warp per brick on the GPU (lanes split the node pairs and the depth
// pixels, reductions by shuffle) and with a single "lane" on the host (tests/hostshim).
struct SerialCtx {
    DFB_HD int lane() const { return 0; }
    DFB_HD int nlanes() const { return 1; }
    DFB_HD float rmin(float v) const { return v; }
    DFB_HD float rmax(float v) const { return v; }
    DFB_HD bool any(bool b) const { return b; }
};
#if defined(__CUDACC__)
// G consecutive lanes cooperate on one brick (G = 8: four bricks per warp; the per-brick work that is uniform across
// lanes -- intervals, projection -- is then amortised over four bricks instead of one).
template <int G>
struct GroupCtx {
    __device__ __forceinline__ unsigned mask() const { return (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) & ~(G - 1))); }
    __device__ __forceinline__ int lane() const { return threadIdx.x & (G - 1); }
    __device__ __forceinline__ int nlanes() const { return G; }
    __device__ __forceinline__ float rmin(float v) const {
        const unsigned m = mask();
        for (int o = G / 2; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(m, v, o));
        return v;
    }
    __device__ __forceinline__ float rmax(float v) const {
        const unsigned m = mask();
        for (int o = G / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(m, v, o));
        return v;
    }
    __device__ __forceinline__ bool any(bool b) const { return (__ballot_sync(mask(), b) & mask()) != 0u; }
};
typedef GroupCtx<32> WarpCtx;
#endif

// Second half of the classification, shared by bricks and regions: project the warped-space box `bx` into every view
// and compare the depth rectangle it covers with the box's depth range, view by view.  Control flow is uniform across `ctx`.
// view_mask: only these views are tested (the others are reported as neither clamp / frus / mixed).
// band_out (optional): bit v set when every voxel of the box is certainly INSIDE the truncation band of view v (every candidate pixel
// carries a measurement and -tdist < tl < tdist for all of them) -- such a voxel is updated with its unclamped value, which only the
// exact tier can produce, so the per-voxel fast tier need not look at it.
template <class Ctx>
DFB_HDN BrickClass box_classify_views(const ProjParams& P, const Box3& bx, int max_rect, const Ctx ctx, int view_mask = 0xff,
                                      int* band_out = nullptr) {
    const float c3[3] = {0.5f * (bx.lo[0] + bx.hi[0]), 0.5f * (bx.lo[1] + bx.hi[1]), 0.5f * (bx.lo[2] + bx.hi[2])};
    const float h3[3] = {0.5f * (bx.hi[0] - bx.lo[0]), 0.5f * (bx.hi[1] - bx.lo[1]), 0.5f * (bx.hi[2] - bx.lo[2])};
    int mask = 0, fr = 0, mixed = 0;
    if (!P.k_pinhole) return brick_class_all_mixed(P.n_views);
    const float mt = 1e-3f * P.tdist_f + 2e-6f * P.coord_mag;
    for (int v = 0; v < P.n_views; ++v) {
        if (!((view_mask >> v) & 1)) continue;
        const ViewFast& V = P.vf[v];
        // camera-space box, then a = X/Z, b = Y/Z, (u,v) = K[0:2] * (a, b, 1): dividing X by Z (not K*lpos rows by each
        // other) keeps the interval dependency problem away from the principal-point term
        float xl, xh, yl, yh, lzl, lzh;
        row_interval(V.T, c3, h3, xl, xh);
        row_interval(V.T + 4, c3, h3, yl, yh);
        row_interval(V.T + 8, c3, h3, lzl, lzh);
        if (!(lzl > 1e-3f * (fabsf(lzh) + 1.f))) { mixed |= 1 << v; continue; }   // must be safely in front of the camera
        float al, ah, bl, bh;
        div_interval(xl, xh, lzl, lzh, al, ah);
        div_interval(yl, yh, lzl, lzh, bl, bh);
        float ul = fminf(P.kf[0] * al, P.kf[0] * ah) + fminf(P.kf[1] * bl, P.kf[1] * bh) + P.kf[2];
        float uh = fmaxf(P.kf[0] * al, P.kf[0] * ah) + fmaxf(P.kf[1] * bl, P.kf[1] * bh) + P.kf[2];
        float vl = fminf(P.kf[3] * al, P.kf[3] * ah) + fminf(P.kf[4] * bl, P.kf[4] * bh) + P.kf[5];
        float vh = fmaxf(P.kf[3] * al, P.kf[3] * ah) + fmaxf(P.kf[4] * bl, P.kf[4] * bh) + P.kf[5];
        {
            const float su = 4e-6f * (fabsf(ul) + fabsf(uh) + fabsf(P.kf[2])) + 1e-4f, sv = 4e-6f * (fabsf(vl) + fabsf(vh) + fabsf(P.kf[5])) + 1e-4f;
            ul -= su; uh += su; vl -= sv; vh += sv;
        }
        const float umax = (float)(P.cols - 1), vmax = (float)(P.rows - 1);
        if (uh < 0.f || ul >= umax || vh < 0.f || vl >= vmax) continue;        // certainly outside this image: view skipped
        if (!(ul >= 0.f && uh < umax && vl >= 0.f && vh < vmax)) { mixed |= 1 << v; continue; }
        const int frbit = 1 << v;
        const int iu0 = (int)rintf(ul), iu1 = (int)rintf(uh), iv0 = (int)rintf(vl), iv1 = (int)rintf(vh);
        const int nu = iu1 - iu0 + 1, npx = nu * (iv1 - iv0 + 1);
        if (npx > max_rect) { mixed |= 1 << v; continue; }
        float zmin = 3.0e38f, zmax = -3.0e38f;
        bool nan = false;
        for (int t = ctx.lane(); t < npx; t += ctx.nlanes()) {
            const int iv = iv0 + t / nu, iu = iu0 + t % nu;
            const float z = -P.depth[v][(size_t)iv * P.cols + iu];
            nan |= !(fabsf(z) <= 3.0e38f);   // NaN or +-inf: the reference's K^-1 product turns an infinite depth into NaN (0 * inf)
            zmin = fminf(zmin, z);
            zmax = fmaxf(zmax, z);
        }
        if (ctx.any(nan)) { mixed |= 1 << v; continue; }
        zmin = ctx.rmin(zmin);
        zmax = ctx.rmax(zmax);
        // kz = Kinv20*u + Kinv21*v + Kinv22 over the rectangle
        const float k0l = fminf(P.kin[0] * ul, P.kin[0] * uh), k0h = fmaxf(P.kin[0] * ul, P.kin[0] * uh);
        const float k1l = fminf(P.kin[1] * vl, P.kin[1] * vh), k1h = fmaxf(P.kin[1] * vl, P.kin[1] * vh);
        const float kzl = k0l + k1l + P.kin[2] - 1e-6f * (fabsf(k0l) + fabsf(k1l) + fabsf(P.kin[2]));
        const float kzh = k0h + k1h + P.kin[2] + 1e-6f * (fabsf(k0h) + fabsf(k1h) + fabsf(P.kin[2]));
        if (!(kzl > 0.f)) { mixed |= 1 << v; continue; }
        if (zmax <= 0.f) { fr |= frbit; continue; }                                // no measurement anywhere: view skipped
        const float zs = 1e-6f * fabsf(zmax) * kzh;
        if (zmin > 0.f && zmin * kzl - lzh > P.tdist_f + mt + zs) { mask |= 1 << v; fr |= frbit; continue; }
        if (zmax * kzh - lzl < -P.tdist_f - mt - zs) { fr |= frbit; continue; }   // every measured pixel lies far in front: skipped
        if (band_out && zmin > 0.f && zmax * kzh - lzl < P.tdist_f - mt - zs && zmin * kzl - lzh > -P.tdist_f + mt + zs) *band_out |= 1 << v;
        mixed |= 1 << v;
    }
    BrickClass r;
    r.clamp = mask; r.frus = fr; r.mixed = mixed;
    r.cls = mixed ? BRICK_CLS_MIXED : mask;
    return r;
}

// rr[15] of a region record: 0 = no bound (fall back to the pairwise hull), 1 = (P_ref, D_R) valid, >= 2: the WHOLE region
// is already classified, code - 2 = CLAMP mask + 256 * frustum bits (every brick of it inherits the class).
DFB_HD float region_code(bool valid, const BrickClass& c) {
    if (!valid) return 0.f;
    if (c.mixed) return 1.f;
    return (float)(2 + c.clamp + 256 * c.frus);
}

// warped-space box of the voxel box (c, h) under the region's reference map, inflated by its deviation bound
DFB_HD void region_box(const float* rr, const float* c, const float* h, float coord_mag, Box3& bx) {
    for (int r = 0; r < 3; ++r) { bx.lo[r] = 3.0e38f; bx.hi[r] = -3.0e38f; }
    box_extend_affine(rr, 1.0f, c, h, bx);
    for (int r = 0; r < 3; ++r) {
        const float m = rr[12 + r] + 2e-3f + 2e-6f * coord_mag;
        bx.lo[r] -= m;
        bx.hi[r] += m;
    }
}

// ---- quads: the per-voxel tier's cheap first look at a MIXED brick -------------------------------------------------------------
// A quad = up to four z-consecutive voxels (x, y, z0 .. z0+nz-1), the unit one thread of the update pass owns.  Under the region's
// reference map P_ref and deviation bound D (|p'(x) - P_ref(x)| <= D for every voxel of the region, see above) the camera-space
// position of a voxel is  lpos = M [x, 1] +- d  with  M = T_v P_ref  and  d = |T_v| D  -- affine in z along the quad, so the pixel
// interval of the whole quad follows from its two end voxels (u = X/Z is monotone along a line in front of the camera) and covers
// one to four depth pixels.  Each voxel is then settled against the depth range of those pixels exactly like a brick is against
// its rectangle: SKIP / CLAMP, certainly INSIDE the band (-> straight to the exact tier), or open (-> pointwise DQB tier).
struct QuadView {
    float M[12];   // lpos = M [x, 1]
    float d[3];    // bound on |lpos - M [x, 1]| per component, rounding of the evaluation included
};
enum { QV_SKIP = 0, QV_CLAMP = 1, QV_BAND = 2, QV_OPEN = 3 };
constexpr int QUAD_MAX_SIDE = 4;   // the quad's pixel rectangle is scanned when it is at most 4 x 4

DFB_HD void quad_view_setup(const ProjParams& P, const float* rr, int v, QuadView& Q) {
    const float* T = P.vf[v].T;
    float D[3];
    for (int r = 0; r < 3; ++r) D[r] = rr[12 + r] + 2e-3f + 2e-6f * P.coord_mag;   // the inflation region_box applies
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 4; ++c) Q.M[4 * r + c] = T[4 * r] * rr[c] + T[4 * r + 1] * rr[4 + c] + T[4 * r + 2] * rr[8 + c] + (c == 3 ? T[4 * r + 3] : 0.f);
        // + float32 rounding of the composed map and of its evaluation at coordinates up to coord_mag
        Q.d[r] = fabsf(T[4 * r]) * D[0] + fabsf(T[4 * r + 1]) * D[1] + fabsf(T[4 * r + 2]) * D[2] + 8e-6f * P.coord_mag + 1e-3f;
    }
}

// States of the quad's voxels for ONE view, 2 bits each (voxel q in bits 2q, 2q+1); *frus = 1 when the quad certainly projects
// inside the image (0: certainly outside when every state is QV_SKIP, unknown when open).
DFB_HD uint32_t quad_view_test(const ProjParams& P, const QuadView& Q, const float* depth, int x, int y, int z0, int nz, int* frus) {
    constexpr uint32_t ALL_OPEN = 0xffu, ALL_SKIP = 0u;
    *frus = 0;
    const float fx = (float)x, fy = (float)y, fz = (float)z0, span = (float)(nz - 1);
    const float X0 = Q.M[0] * fx + Q.M[1] * fy + Q.M[2] * fz + Q.M[3];
    const float Y0 = Q.M[4] * fx + Q.M[5] * fy + Q.M[6] * fz + Q.M[7];
    const float Z0 = Q.M[8] * fx + Q.M[9] * fy + Q.M[10] * fz + Q.M[11];
    const float X3 = X0 + span * Q.M[2], Y3 = Y0 + span * Q.M[6], Z3 = Z0 + span * Q.M[10];
    const float zlo = fminf(Z0, Z3) - Q.d[2];
    if (!(zlo > 1e-3f * (fmaxf(fabsf(Z0), fabsf(Z3)) + Q.d[2] + 1.f))) return ALL_OPEN;   // must be safely in front of the camera
    const float r0 = fast_rcp(Z0), r3 = fast_rcp(Z3);
    const float a0 = X0 * r0, a3 = X3 * r3, b0 = Y0 * r0, b3 = Y3 * r3;
    const float rl = fast_rcp(zlo) * 1.00001f;                                 // >= 1 / (true lpos_z) anywhere on the quad
    const float am = fmaxf(fabsf(a0), fabsf(a3)), bm = fmaxf(fabsf(b0), fabsf(b3));
    // |X'/Z' - X/Z| <= (dX + |X/Z| dZ) / Z'   for |X' - X| <= dX, |Z' - Z| <= dZ
    const float ea = (Q.d[0] + am * Q.d[2]) * rl + 4e-6f * am + 1e-6f, eb = (Q.d[1] + bm * Q.d[2]) * rl + 4e-6f * bm + 1e-6f;
    const float al = fminf(a0, a3) - ea, ah = fmaxf(a0, a3) + ea, bl = fminf(b0, b3) - eb, bh = fmaxf(b0, b3) + eb;
    // plain pinhole (no skew, K^-1 row 2 = (0,0,1): the usual camera) keeps two multiplies per axis; the general form is the interval
    // product of the brick classifier
    const bool plain = P.kf[1] == 0.f && P.kf[3] == 0.f && P.kf[0] > 0.f && P.kf[4] > 0.f && P.kin[0] == 0.f && P.kin[1] == 0.f;
    float ul, uh, vl, vh;
    if (plain) {
        ul = P.kf[0] * al + P.kf[2]; uh = P.kf[0] * ah + P.kf[2];
        vl = P.kf[4] * bl + P.kf[5]; vh = P.kf[4] * bh + P.kf[5];
    } else {
        ul = fminf(P.kf[0] * al, P.kf[0] * ah) + fminf(P.kf[1] * bl, P.kf[1] * bh) + P.kf[2];
        uh = fmaxf(P.kf[0] * al, P.kf[0] * ah) + fmaxf(P.kf[1] * bl, P.kf[1] * bh) + P.kf[2];
        vl = fminf(P.kf[3] * al, P.kf[3] * ah) + fminf(P.kf[4] * bl, P.kf[4] * bh) + P.kf[5];
        vh = fmaxf(P.kf[3] * al, P.kf[3] * ah) + fmaxf(P.kf[4] * bl, P.kf[4] * bh) + P.kf[5];
    }
    {
        const float su = 4e-6f * (fabsf(ul) + fabsf(uh) + fabsf(P.kf[2])) + 1e-4f, sv = 4e-6f * (fabsf(vl) + fabsf(vh) + fabsf(P.kf[5])) + 1e-4f;
        ul -= su; uh += su; vl -= sv; vh += sv;
    }
    const float umax = (float)(P.cols - 1), vmax = (float)(P.rows - 1);
    if (uh < 0.f || ul >= umax || vh < 0.f || vl >= vmax) return ALL_SKIP;           // certainly outside this image
    if (!(ul >= 0.f && uh < umax && vl >= 0.f && vh < vmax)) return ALL_OPEN;
    const int iu0 = (int)rintf(ul), iu1 = (int)rintf(uh), iv0 = (int)rintf(vl), iv1 = (int)rintf(vh);
    if (iu1 - iu0 >= QUAD_MAX_SIDE || iv1 - iv0 >= QUAD_MAX_SIDE) return ALL_OPEN;
    float zmin = 3.0e38f, zmax = -3.0e38f;
    bool bad = false;
    {
        // the rectangle is at most QUAD_MAX_SIDE x QUAD_MAX_SIDE: fixed trip counts, pixels beyond its far edges re-read the edge
        const float* row = depth + (size_t)iv0 * P.cols + iu0;
        const int nu = iu1 - iu0, nv = iv1 - iv0;
#pragma unroll
        for (int dv = 0; dv < QUAD_MAX_SIDE; ++dv) {
            if (dv > nv) break;
#pragma unroll
            for (int du = 0; du < QUAD_MAX_SIDE; ++du) {
                if (du > nu) break;
                const float z = -row[du];
                bad |= !(fabsf(z) <= 3.0e38f);
                zmin = fminf(zmin, z);
                zmax = fmaxf(zmax, z);
            }
            row += P.cols;
        }
    }
    if (bad) return ALL_OPEN;
    float kzl = 1.f, kzh = 1.f;
    if (!plain) {
        const float k0l = fminf(P.kin[0] * ul, P.kin[0] * uh), k0h = fmaxf(P.kin[0] * ul, P.kin[0] * uh);
        const float k1l = fminf(P.kin[1] * vl, P.kin[1] * vh), k1h = fmaxf(P.kin[1] * vl, P.kin[1] * vh);
        kzl = k0l + k1l + P.kin[2] - 1e-6f * (fabsf(k0l) + fabsf(k1l) + fabsf(P.kin[2]));
        kzh = k0h + k1h + P.kin[2] + 1e-6f * (fabsf(k0h) + fabsf(k1h) + fabsf(P.kin[2]));
    } else {
        kzl = P.kin[2] - 1e-6f * fabsf(P.kin[2]);
        kzh = P.kin[2] + 1e-6f * fabsf(P.kin[2]);
    }
    if (!(kzl > 0.f)) return ALL_OPEN;
    *frus = 1;
    if (zmax <= 0.f) return ALL_SKIP;                                                // no measurement on any candidate pixel
    const float mt = 1e-3f * P.tdist_f + 2e-6f * P.coord_mag, zs = 1e-6f * fabsf(zmax) * kzh;
    const float thr = P.tdist_f + mt + zs, inb = P.tdist_f - mt - zs;
    const float cl = zmin * kzl, ch = zmax * kzh;
    uint32_t st = 0;
    for (int q = 0; q < nz; ++q) {
        const float Zq = Z0 + (float)q * Q.M[10];
        const float lzl = Zq - Q.d[2], lzh = Zq + Q.d[2];
        uint32_t s;
        if (zmin > 0.f && cl - lzh > thr) s = QV_CLAMP;
        else if (ch - lzl < -thr) s = QV_SKIP;
        else if (zmin > 0.f && ch - lzl < inb && cl - lzh > -inb) s = QV_BAND;
        else s = QV_OPEN;
        st |= s << (2 * q);
    }
    return st;
}

// All views of a quad.  qv: the brick's QuadView per view; views / m0 / f0: the views the brick's box test left open and the bits
// of the views it settled.  Output per voxel q: state[q] (QV_SKIP = settled: then m[q] / f[q] are its clamp / frustum bits of all
// views; QV_BAND: defer to the exact tier; QV_OPEN: pointwise tier).
DFB_HD void quad_pretest(const ProjParams& P, const QuadView* qv, int x, int y, int z0, int nz, int views, int m0, int f0, int* state, int* m, int* f) {
    for (int q = 0; q < 4; ++q) { state[q] = QV_SKIP; m[q] = m0; f[q] = f0; }
    if (!P.k_pinhole) {   // u = (K lpos)_0 / (K lpos)_2 with a general third row: the X/Z, Y/Z intervals below do not apply
        for (int q = 0; q < nz; ++q) state[q] = QV_OPEN;
        return;
    }
    for (int v = 0; v < P.n_views; ++v) {
        if (!((views >> v) & 1)) continue;
        int fr;
        const uint32_t st = quad_view_test(P, qv[v], P.depth[v], x, y, z0, nz, &fr);
        for (int q = 0; q < nz; ++q) {
            const int s = (int)((st >> (2 * q)) & 3u);
            if (s == QV_CLAMP) m[q] |= 1 << v;
            if (s == QV_BAND) state[q] = QV_BAND;
            else if (s == QV_OPEN && state[q] != QV_BAND) state[q] = QV_OPEN;
            if (fr) f[q] |= 1 << v;
        }
    }
}

// Classify brick (bxs,by,bz) (bxs slab-local).  All control flow is uniform across the lanes of `ctx`.
template <class Ctx>
DFB_HDN BrickClass brick_classify(const ProjParams& P, const uint16_t* brick_nodes, const uint8_t* brick_count, const uint32_t* brick_pairs,
                                  const float* region_rec, int nby, int nbz, int bxs, int by, int bz, const Ctx ctx) {
    const int xlo = P.x0 + bxs * BRICK_X, ylo = by * BRICK_Y, zlo = bz * BRICK_Z;
    const int xhi = (xlo + BRICK_X - 1 < P.x1 - 1) ? xlo + BRICK_X - 1 : P.x1 - 1;
    const int yhi = (ylo + BRICK_Y - 1 < P.ry - 1) ? ylo + BRICK_Y - 1 : P.ry - 1;
    const int zhi = (zlo + BRICK_Z - 1 < P.rz - 1) ? zlo + BRICK_Z - 1 : P.rz - 1;
    const float c[3] = {0.5f * (xlo + xhi), 0.5f * (ylo + yhi), 0.5f * (zlo + zhi)};
    const float h[3] = {0.5f * (xhi - xlo), 0.5f * (yhi - ylo), 0.5f * (zhi - zlo)};
    Box3 bx;
    bool have_box = false;
    if (!P.rigid && region_rec) {
        // O(1) path: the region's reference affine map applied to the brick, inflated by the region's deviation bound
        const int nry = (P.ry + REGION_Y - 1) / REGION_Y, nrz = (P.rz + REGION_Z - 1) / REGION_Z;
        const float* rr = region_rec + (((size_t)(bxs * BRICK_X / REGION_X) * nry + by * BRICK_Y / REGION_Y) * nrz + bz * BRICK_Z / REGION_Z) * REGION_REC_FLOATS;
        if (rr[15] > 1.5f) {           // the region as a whole is already SKIP / CLAMP
            const int code = (int)rr[15] - 2;
            BrickClass r;
            r.clamp = code & 0xff; r.frus = code >> 8; r.mixed = 0; r.cls = r.clamp;
            return r;
        }
        if (rr[15] > 0.5f) {
            region_box(rr, c, h, P.coord_mag, bx);
            have_box = true;
        }
    }
    if (have_box) {
    } else if (P.rigid) {
        for (int r = 0; r < 3; ++r) { bx.lo[r] = c[r] - h[r]; bx.hi[r] = c[r] + h[r]; }
    } else {
        const size_t b = ((size_t)bxs * nby + by) * nbz + bz;
        const int cnt = brick_count[b];
        if (cnt == 0 || cnt > BRICK_MAXC) return brick_class_all_mixed(P.n_views);
        for (int r = 0; r < 3; ++r) { bx.lo[r] = 3.0e38f; bx.hi[r] = -3.0e38f; }
        const uint16_t* ids = brick_nodes + b * BRICK_MAXC;
        bool bad = false;
        const int npairs_all = cnt * (cnt + 1) / 2;
        // pairs to visit: the set bits of the cached co-occurrence mask (all pairs when no mask is given), dealt out to
        // the lanes in compact order so that no lane idles on pairs that never blend inside this brick
        uint32_t pm[BRICK_PAIR_WORDS];
        int pre[BRICK_PAIR_WORDS + 1];
        pre[0] = 0;
        for (int w = 0; w < BRICK_PAIR_WORDS; ++w) {
            uint32_t m = brick_pairs ? brick_pairs[b * BRICK_PAIR_WORDS + w] : 0xffffffffu;
            const int rem = npairs_all - 32 * w;
            if (rem <= 0) m = 0u;
            else if (rem < 32) m &= (1u << rem) - 1u;
            pm[w] = m;
            pre[w + 1] = pre[w] + popc32(m);
        }
        const int npairs = pre[BRICK_PAIR_WORDS];
        for (int t = ctx.lane(); t < npairs; t += ctx.nlanes()) {
            int w = 0;
            while (pre[w + 1] <= t) ++w;
            const int p = 32 * w + nth_set_bit(pm[w], t - pre[w]);
            // p -> (i, j), j <= i, row-major lower triangle
            int i = (int)((sqrtf(8.f * (float)p + 1.f) - 1.f) * 0.5f);
            while (i * (i + 1) / 2 > p) --i;
            while ((i + 1) * (i + 2) / 2 <= p) ++i;
            const int j = p - i * (i + 1) / 2;
            const float4 r1 = P.node_rec[3 * (size_t)ids[i] + 1];
            const float4 r2 = P.node_rec[3 * (size_t)ids[i] + 2];
            const float qi[8] = {r1.x, r1.y, r1.z, r1.w, r2.x, r2.y, r2.z, r2.w};
            float nii = 0.f;
            for (int t = 0; t < 8; ++t) nii += qi[t] * qi[t];
            float Ap[12];
            if (j == i) {
                // every blend weight must be a normal float32 in the reference (`w * dg_dq` rounds to float32; weights that
                // vanish can make the blend fall back to the identity): test the farthest voxel of the brick
                const float4 r0 = P.node_rec[3 * (size_t)ids[i]];
                const float ddx = fabsf(c[0] - r0.x) + h[0], ddy = fabsf(c[1] - r0.y) + h[1], ddz = fabsf(c[2] - r0.z) + h[2];
                if (!((ddx * ddx + ddy * ddy + ddz * ddz) * r0.w > -125.f) || !(nii > 1e-20f)) { bad = true; continue; }
                dq_affine_polar2(qi, qi, Ap);
                box_extend_affine(Ap, 0.5f / nii, c, h, bx);
            } else {
                const float4 s1 = P.node_rec[3 * (size_t)ids[j] + 1];
                const float4 s2 = P.node_rec[3 * (size_t)ids[j] + 2];
                const float qj[8] = {s1.x, s1.y, s1.z, s1.w, s2.x, s2.y, s2.z, s2.w};
                float ip = 0.f, njj = 0.f;
                for (int t = 0; t < 8; ++t) { ip += qi[t] * qj[t]; njj += qj[t] * qj[t]; }
                // <q_i,q_j> must be safely positive (else the convex-combination argument does not apply)
                if (!(ip > 0.25f * sqrtf(nii * njj))) { bad = true; continue; }
                dq_affine_polar2(qi, qj, Ap);
                box_extend_affine(Ap, 0.5f / ip, c, h, bx);
            }
        }
        if (ctx.any(bad)) return brick_class_all_mixed(P.n_views);
        // reference-side rounding of p' (Q3: float32 cast) and float64 noise; polarisation cancellation slack
        for (int r = 0; r < 3; ++r) {
            const float m = 2e-3f + 2e-6f * P.coord_mag;
            bx.lo[r] = ctx.rmin(bx.lo[r]) - m;
            bx.hi[r] = ctx.rmax(bx.hi[r]) + m;
        }
    }
    return box_classify_views(P, bx, BRICK_MAX_RECT, ctx);
}

}  // namespace dfb
