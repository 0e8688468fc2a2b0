// dfb_voxel.h -- per-voxel update logic of the three TSDF passes (a1 volume-sampling, a2 rigid
// projective, a3 warped projective), shared by the CUDA kernels (tsdf.cu) and the CPU logic tests
// (tests/hostshim/).  See dfb_math.h for the two arithmetic tiers.
#pragma once
#include "dfb_math.h"

namespace dfb {

enum { CLS_SKIP = 0, CLS_CLAMP = 1, CLS_UNCERTAIN = 2 };

struct ViewFast {
    float P[12];  // K * E_v * A_lw   (rows: u*z, v*z, z_k)
    float L[4];   // third row of E_v * A_lw  (lpos_z)
    float T[12];  // E_v * A_lw  (lpos), used by the brick classifier
};

struct ProjParams {
    // volume (slab)
    float* tsdf;
    float* weight;
    int rx, ry, rz, x0, x1;
    // warp field
    const float4* node_rec;
    const float* node_pos;
    const float* node_dq;
    const float* node_w;
    const uint16_t* knn;
    int k;
    int has_lw, lw_is_f32;
    double lw[8];
    // a2 only: grid -> world (pos = scale*(idx - res/2) + center), rigid 3x4 lw
    int rigid;
    double g_scale, g_half, g_center[3];
    double lw34[12];
    float G[12];  // fp32 affine idx -> lpos (a2 fast tier)
    // views
    int n_views, rows, cols, has_E;
    const float* depth[8];
    double K[9], Kinv[9];
    double E[8][12];
    ViewFast vf[8];
    float kf[6];     // K rows 0,1 in fp32
    int k_pinhole;   // K row 2 == (0,0,1): u = K00*X/Z + K01*Y/Z + K02 (brick classifier requirement)
    float kin[3];    // Kinv row 2, fp32
    float tnorm;     // max abs row sum of the view transforms E_v * A_lw (3x3 parts): lpos error per unit p' error
    float knorm;     // max abs row sum of K rows 0,1 (error propagation to pixels)
    float ezf;       // abs row sum of K row 2 (1 for a pinhole), never below 1: error propagation to the projective divisor
    float kin_uv;    // |Kinv20| + |Kinv21|
    float coord_mag; // bound on |coordinates| entering the fp32 chain (error scale)
    // update
    double tdist, wmax, scale;
    float tdist_f, wmax_f;
    // bookkeeping
    uint32_t* list;
    uint32_t capacity;
    uint32_t* counters;
    uint32_t* overflow_bits;   // optional: deferred voxels that did not fit the list (one bit per slab voxel)
    uint8_t* mask_out;
    uint8_t* frustum_out;
};

struct VolParams {
    float* tsdf;
    float* weight;
    int rx, ry, rz, x0, x1;
    const float4* node_rec;
    const float* node_pos;
    const float* node_dq;
    const float* node_w;
    const uint16_t* knn;
    int k;
    int has_lw, lw_is_f32;
    double lw[8];
    float A[12];  // fp32 affine of lw (identity when !has_lw)
    const float* curr;
    int cx, cy, cz;
    double tdist, wmax;
    float tdist_f, wmax_f, coord_mag;
    uint32_t* list;
    uint32_t capacity;
    uint32_t* counters;
    uint32_t* overflow_bits;
    uint8_t* mask_out;
};

// Where the packed node records are read from: global memory (any caller), or the shared-memory copy of the whole table that
// the production update kernel keeps (explicit ld.shared: a generic pointer would cost the address-space check on every gather).
struct RecGlobal {
    const float4* p;
    DFB_HD float4 ld(size_t i) const { return p[i]; }
};
#if defined(__CUDACC__)
struct RecShared {
    uint32_t base;   // shared-window address of record 0
    __device__ __forceinline__ float4 ld(size_t i) const {
        float4 r;
        asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(base + (uint32_t)i * 16u));
        return r;
    }
};
#endif

// ---- fast tier: DQB of k nodes in fp32 --------------------------------------------------------
// rec layout: [pos.xyz, coef][dq0..3][dq4..7], coef = -log2(e)/(4 w^2)  (exp(-(d/2w)^2) = exp2(coef*d^2))
// Returns false when the fp32 chain cannot be trusted at all (all weights underflow in the reference's
// float64 exp, or zero blend) -> caller treats the voxel as uncertain.
// amin_out: most negative exp argument (natural-log units) among the k nodes: scales the error bound.
template <int KMAX, class Rec>
DFB_HD bool blend_warp_fast(const Rec rec, const uint16_t* ids, int k_rt, float px, float py, float pz, float* out,
                            float* amin_out, bool exact_k = false) {
    const int k = exact_k ? KMAX : k_rt;   // exact_k is a compile-time constant at every call site: predicates fold away
    float a[KMAX];
    float amax = -3.0e38f, amin = 0.f;
#pragma unroll
    for (int i = 0; i < KMAX; ++i) {
        if (i < k) {
            const float4 r0 = rec.ld(3 * (size_t)ids[i]);
            const float dx = px - r0.x, dy = py - r0.y, dz = pz - r0.z;
            a[i] = (dx * dx + dy * dy + dz * dz) * r0.w;
            amax = fmaxf(amax, a[i]);
            amin = fminf(amin, a[i]);
        }
    }
    *amin_out = amin * 0.69314718f;
    // The reference multiplies each weight into the float32 node dq as a float32 (`w * dg_dq`, python float is weak):
    // weights below the float32 normal range (exp arg < -87.3, i.e. < -126 in log2 units) lose precision or vanish, and
    // a blend whose weights all vanish falls back to the identity (core/fusion.py:544-549).  The fp32 tier only handles
    // voxels whose LARGEST weight is a normal float32; the rest is decided by the exact tier.
    if (!(amax > -125.f)) return false;
    float b[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < KMAX; ++i) {
        if (i < k) {
            const float w = fast_exp2(a[i] - amax);
            const float4 r1 = rec.ld(3 * (size_t)ids[i] + 1);
            const float4 r2 = rec.ld(3 * (size_t)ids[i] + 2);
            b[0] += w * r1.x; b[1] += w * r1.y; b[2] += w * r1.z; b[3] += w * r1.w;
            b[4] += w * r2.x; b[5] += w * r2.y; b[6] += w * r2.z; b[7] += w * r2.w;
        }
    }
    float n2 = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) n2 += b[c] * b[c];
    if (!(n2 > 1e-30f)) return false;
    float o[3];
    dq_apply_unnormalised(b, px, py, pz, o);
    const float inv = fast_rcp(n2);
    out[0] = o[0] * inv; out[1] = o[1] * inv; out[2] = o[2] * inv;
    return true;
}

// error bound (voxel units) of the fp32 warp chain against the reference's value
DFB_HD float pos_err_bound(float coord_mag, float amin_nat) {
    return 1.1920929e-7f * coord_mag * (32.f - 4.f * amin_nat);
}

// ---- fast tier: classify one view (a2/a3) -----------------------------------------------------
// in: (pu,pv,pz) = K*lpos, lz = lpos_z, e = error bound on lpos components.
// frustum_bit: 1 if certainly inside the image, 0 if certainly outside (only valid if result != UNCERTAIN)
DFB_HD int classify_view(float pu, float pv, float pz, float lz, float e, const float* depth, int rows, int cols,
                         const float* kin, float knorm, float ezf, float kin_uv, float tdist, int* frustum_bit) {
    *frustum_bit = 0;
    const float apz = fabsf(pz);
    const float ez = e * ezf;  // error of the divisor K_row2 . lpos
    if (!(apz > 64.f * ez)) return CLS_UNCERTAIN;  // also catches NaN
    const float inv = fast_rcp(pz);
    const float u = pu * inv, v = pv * inv;
    const float ainv = fabsf(inv) * 1.02f;         // 1/|true divisor| <= 1/(|pz| - ez) <= (1/|pz|) / (1 - 1/64)
    const float eu = (knorm * e + fabsf(u) * ez) * ainv + 4.8e-7f * fabsf(u) + 1e-6f;
    const float ev = (knorm * e + fabsf(v) * ez) * ainv + 4.8e-7f * fabsf(v) + 1e-6f;
    const float umax = (float)(cols - 1), vmax = (float)(rows - 1);
    // certainly outside?
    if (u < -eu || u >= umax + eu || v < -ev || v >= vmax + ev) return CLS_SKIP;
    // certainly inside?
    if (!(u >= eu && u < umax - eu && v >= ev && v < vmax - ev)) return CLS_UNCERTAIN;
    *frustum_bit = 1;
    // candidate pixels (round-half-even both ends of the uncertainty interval)
    const int u0 = (int)rintf(u - eu), u1 = (int)rintf(u + eu);
    const int v0 = (int)rintf(v - ev), v1 = (int)rintf(v + ev);
    const float kz = kin[0] * u + kin[1] * v + kin[2];
    const float et0 = kin_uv * (eu + ev) + 4.8e-7f * fabsf(kz);
    const float et1 = e + 2.4e-7f * fabsf(lz);
    // The rounding of (u,v) is normally unambiguous (one candidate pixel).  When the uncertainty interval straddles a
    // rounding boundary in ONE axis there are two candidates; both are sampled, branch-free, and must agree.  Anything
    // wider (both axes ambiguous, or an interval longer than a pixel) is left to the exact tier.
    if ((u0 != u1 && v0 != v1) || u1 - u0 > 1 || v1 - v0 > 1) return CLS_UNCERTAIN;
    const float za = -depth[(size_t)v0 * cols + u0];
    const float zb = -depth[(size_t)v1 * cols + u1];
    const float tla = za * kz - lz, tlb = zb * kz - lz;
    const float eta = fabsf(za) * et0 + et1, etb = fabsf(zb) * et0 + et1;
    const bool skip_a = !(za > 0.f) || tla < -tdist - eta, skip_b = !(zb > 0.f) || tlb < -tdist - etb;
    const bool clamp_a = za > 0.f && tla > tdist + eta, clamp_b = zb > 0.f && tlb > tdist + etb;
    if (skip_a && skip_b) return CLS_SKIP;
    if (clamp_a && clamp_b) return CLS_CLAMP;
    return CLS_UNCERTAIN;
}

// ---- exact tier: whole-voxel functions --------------------------------------------------------
// a3 / a2 for one voxel; v,w in/out (float32 storage, float64 arithmetic).  Returns mask bits in
// *mask, frustum bits in *frus.
template <int KT = 0>
DFB_HDN void voxel_projective_exact(const ProjParams& P, int x, int y, int z, const uint16_t* ids16, float* v_io,
                                    float* w_io, int* mask, int* frus) {
    double base[3];
    if (P.rigid) {
        // pos = scale * (pos - sdf_center) + center  (core/fusion_dm.py:183,191), float64
        const double idx[3] = {(double)x, (double)y, (double)z};
        for (int c = 0; c < 3; ++c) base[c] = dadd(dmul(P.g_scale, dsub(idx[c], P.g_half)), P.g_center[c]);
    } else {
        const float p[3] = {(float)x, (float)y, (float)z};
        int ids[8];
        if (KT > 0) {
#pragma unroll
            for (int i = 0; i < (KT > 0 ? KT : 1); ++i) ids[i] = ids16[i];
        } else {
            for (int i = 0; i < P.k; ++i) ids[i] = ids16[i];
        }
        // -DDFB_EXACT_REC=1: fetch pos / dq from the packed records (3 x 16 B per node) instead of the SoA arrays (11 scalar loads);
        // measured at 512^3: 0.160 ms against 0.149 ms -- the pass is not load-bound, the wider loads only lengthen the dependency
        // chains -- so the scalar form stays the default
#ifndef DFB_EXACT_REC
#define DFB_EXACT_REC 0
#endif
        warp_ref<KT>(p, nullptr, ids, P.k, P.node_pos, P.node_dq, P.node_w, P.lw, P.has_lw != 0, P.lw_is_f32 != 0, base,
                     nullptr, nullptr, true, DFB_EXACT_REC ? P.node_rec : nullptr);
    }
    double v = (double)*v_io, w = (double)*w_io;
    int m = 0, f = 0;
    for (int vi = 0; vi < P.n_views; ++vi) {
        double lpos[3];
        if (P.rigid) {
            for (int r = 0; r < 3; ++r) lpos[r] = dot4_ref(P.lw34 + 4 * r, base[0], base[1], base[2], 1.0);
        } else if (P.has_E) {
            for (int r = 0; r < 3; ++r) lpos[r] = dot4_ref(P.E[vi] + 4 * r, base[0], base[1], base[2], 1.0);
        } else {
            lpos[0] = base[0]; lpos[1] = base[1]; lpos[2] = base[2];
        }
        const int r = project_fuse_ref(lpos, P.depth[vi], P.rows, P.cols, P.K, P.Kinv, P.tdist, P.scale, P.wmax, &v, &w);
        if (r & 1) m |= 1 << vi;
        if (r & 2) f |= 1 << vi;
    }
    *v_io = (float)v;
    *w_io = (float)w;
    *mask = m;
    *frus = f;
}

// fast classification of a voxel for a2/a3.  Returns CLS_UNCERTAIN, or CLS_SKIP/CLS_CLAMP-style result:
// *mask gets the per-view clamp bits (all certain), *frus the per-view frustum bits.
// ONEVIEW: n_views == 1 known at compile time (the view record is then addressed with immediate offsets)
// views / m0 / f0: a MIXED brick hands over the views its box test left open (bit mask) and the CLAMP / frustum bits of the
// views it settled; only the open views are evaluated here.  Defaults = every view open.
template <int KMAX, bool EXACTK, bool ONEVIEW, class Rec>
DFB_HD int voxel_projective_classify_rec(const ProjParams& P, int x, int y, int z, const uint16_t* ids, int* mask, int* frus,
                                         int views, int m0, int f0, const Rec rec) {
    float pw[3];
    float e;
    if (P.rigid) {
        pw[0] = (float)x; pw[1] = (float)y; pw[2] = (float)z;
        e = 1.1920929e-7f * P.coord_mag * 32.f;
    } else {
        float amin;
        if (!blend_warp_fast<KMAX>(rec, ids, P.k, (float)x, (float)y, (float)z, pw, &amin, EXACTK)) return CLS_UNCERTAIN;
        e = pos_err_bound(P.coord_mag, amin);
    }
    int m = m0, f = f0;
    const int nv = ONEVIEW ? 1 : P.n_views;
#pragma unroll 1
    for (int vi = 0; vi < nv; ++vi) {
        if (!ONEVIEW && !((views >> vi) & 1)) continue;
        const ViewFast& V = P.vf[ONEVIEW ? 0 : vi];
        const float pu = V.P[0] * pw[0] + V.P[1] * pw[1] + V.P[2] * pw[2] + V.P[3];
        const float pv = V.P[4] * pw[0] + V.P[5] * pw[1] + V.P[6] * pw[2] + V.P[7];
        const float pz = V.P[8] * pw[0] + V.P[9] * pw[1] + V.P[10] * pw[2] + V.P[11];
        const float lz = V.L[0] * pw[0] + V.L[1] * pw[1] + V.L[2] * pw[2] + V.L[3];
        int fb;
        const int c = classify_view(pu, pv, pz, lz, e, P.depth[ONEVIEW ? 0 : vi], P.rows, P.cols, P.kin, P.knorm, P.ezf, P.kin_uv, P.tdist_f, &fb);
        if (c == CLS_UNCERTAIN) return CLS_UNCERTAIN;
        if (c == CLS_CLAMP) m |= 1 << vi;
        if (fb) f |= 1 << vi;
    }
    *mask = m;
    *frus = f;
    return m ? CLS_CLAMP : CLS_SKIP;
}

template <int KMAX, bool EXACTK = false, bool ONEVIEW = false>
DFB_HD int voxel_projective_classify(const ProjParams& P, int x, int y, int z, const uint16_t* ids, int* mask, int* frus,
                                     int views = 0xff, int m0 = 0, int f0 = 0) {
    const RecGlobal rec = {P.node_rec};
    return voxel_projective_classify_rec<KMAX, EXACTK, ONEVIEW>(P, x, y, z, ids, mask, frus, views, m0, f0, rec);
}

// a1 for one voxel, exact tier.
template <int KT = 0>   // KT > 0: compile-time neighbour count (== P.k)
DFB_HDN bool voxel_volume_exact(const VolParams& P, int x, int y, int z, const uint16_t* ids16, float* v_io, float* w_io) {
    const float p[3] = {(float)x, (float)y, (float)z};
    int ids[8];
    if (KT > 0) {
#pragma unroll
        for (int i = 0; i < (KT > 0 ? KT : 1); ++i) ids[i] = ids16[i];
    } else {
        for (int i = 0; i < P.k; ++i) ids[i] = ids16[i];
    }
    double pw[3];
    float wi = 0.f;
    warp_ref<KT>(p, nullptr, ids, P.k, P.node_pos, P.node_dq, P.node_w, P.lw, P.has_lw != 0, P.lw_is_f32 != 0, pw, nullptr, &wi, true);
    double tl = 0.0;
    const bool valid = interpolate_tsdf_ref(pw, P.curr, P.cx, P.cy, P.cz, &tl);
    double v = (double)*v_io, w = (double)*w_io;
    const bool upd = volume_fuse_ref(valid, tl, wi, P.k, P.tdist, P.wmax, &v, &w);
    if (upd) { *v_io = (float)v; *w_io = (float)w; }
    return upd;
}

// a1 fast classification: CLS_SKIP (certainly outside the live volume, or every corner below -tdist), CLS_CLAMP (all 8 corners >= tdist:
// the update uses min(tdist, tl) = tdist up to 1e-16), else CLS_UNCERTAIN.  wi_out: Q4 weight (fp32).
template <int KMAX>
DFB_HD int voxel_volume_classify(const VolParams& P, int x, int y, int z, const uint16_t* ids, float* wi_out) {
    float pw[3];
    float e;
    const float px = (float)x, py = (float)y, pz = (float)z;
    if (P.k > 0) {
        float amin;
        const RecGlobal rec = {P.node_rec};
        if (!blend_warp_fast<KMAX>(rec, ids, P.k, px, py, pz, pw, &amin)) return CLS_UNCERTAIN;
        e = pos_err_bound(P.coord_mag, amin);
    } else {
        pw[0] = px; pw[1] = py; pw[2] = pz;
        e = 1.1920929e-7f * P.coord_mag * 32.f;
    }
    const float qx = P.A[0] * pw[0] + P.A[1] * pw[1] + P.A[2] * pw[2] + P.A[3];
    const float qy = P.A[4] * pw[0] + P.A[5] * pw[1] + P.A[6] * pw[2] + P.A[7];
    const float qz = P.A[8] * pw[0] + P.A[9] * pw[1] + P.A[10] * pw[2] + P.A[11];
    const float mx = (float)(P.cx - 1), my = (float)(P.cy - 1), mz = (float)(P.cz - 1);
    if (qx < -e || qy < -e || qz < -e || qx > mx + e || qy > my + e || qz > mz + e) return CLS_SKIP;
    if (!(qx >= e && qy >= e && qz >= e && qx <= mx - e && qy <= my - e && qz <= mz - e)) return CLS_UNCERTAIN;
    // corners of every cell the true position may fall in
    const int x0 = (int)floorf(qx - e), x1 = (int)ceilf(qx + e);
    const int y0 = (int)floorf(qy - e), y1 = (int)ceilf(qy + e);
    const int z0 = (int)floorf(qz - e), z1 = (int)ceilf(qz + e);
    float mn = 3.0e38f, mxc = -3.0e38f;
    for (int a = x0; a <= x1; ++a)
        for (int b = y0; b <= y1; ++b)
            for (int c = z0; c <= z1; ++c) {
                const float t = P.curr[((size_t)a * P.cy + b) * P.cz + c];
                mn = fminf(mn, t);
                mxc = fmaxf(mxc, t);
            }
    // the Q1 interpolation is a convex combination of the corners (its swapped y/z weights are still in [0,1]): with every
    // corner below -tdist the reference's `tsdf_l > -tdist` (core/fusion.py:179) is certainly false
    if (mxc <= -P.tdist_f * 1.000001f) return CLS_SKIP;
    // every corner >= tdist: the update is min(tdist, tl) = tdist.  Equality is included (a live TSDF clipped at +tdist is the common
    // input): the reference's float64 interpolation of corners that all equal tdist can come out an ulp of a double below it, which
    // moves the stored float32 by at most its own last bit -- far inside the 1e-5 tdist value tolerance, and no mask depends on it.
    // (The mirror case at -tdist decides a MASK, so it keeps its margin and goes to the exact tier.)
    if (!(mn >= P.tdist_f)) return CLS_UNCERTAIN;
    if (P.k > 0) {
        float wi = 0.f;
        for (int i = 0; i < P.k; ++i) {
            const float4 r0 = P.node_rec[3 * (size_t)ids[i]];
            const float nr = norm3_f32_ref(r0.x, r0.y, r0.z, px, py, pz);
            const float term = fdiv(nr, (float)P.k);
            wi = (i == 0) ? term : fadd(wi, term);
        }
        *wi_out = wi;
    } else {
        *wi_out = 1.f;
    }
    return CLS_CLAMP;
}

}  // namespace dfb
