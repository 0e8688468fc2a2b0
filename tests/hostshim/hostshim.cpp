// hostshim.cpp -- TEST INFRASTRUCTURE ONLY.  Compiles the per-voxel logic of the CUDA kernels
// (dynamicfusion_body_b200/csrc/dfb_voxel.h, dfb_params.h) for the host so that the CPU test-suite can
// check the fast-tier classification and the reference-exact tier against the oracle on a box that has
// no GPU.  It mirrors the composition the kernels in tsdf.cu perform (fast pass, then exact pass over
// the uncertain voxels) with plain loops and HOST pointers.  The product never loads this library.
#include <stdarg.h>
#include <stdio.h>
#include <algorithm>
#include <array>
#include <vector>

#include "../../dynamicfusion_body_b200/csrc/dfb_params.h"
#include "../../dynamicfusion_body_b200/csrc/dfb_brick.h"

namespace dfb {
static char g_err[512];
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace dfb

using namespace dfb;

extern "C" const char* hs_last_error() { return g_err; }

extern "C" void hs_nodes_pack(const float* pos, const float* dq, const float* w, int n, float* rec) {
    for (int i = 0; i < n; ++i) {
        float* r = rec + (size_t)i * DFB_NODE_REC_FLOATS;
        r[0] = pos[3 * i]; r[1] = pos[3 * i + 1]; r[2] = pos[3 * i + 2];
        const double ww = (double)w[i];
        r[3] = (float)(-1.4426950408889634 / (4.0 * ww * ww));
        for (int c = 0; c < 8; ++c) r[4 + c] = dq[8 * i + c];
    }
}

extern "C" void hs_brick_nodes_build(const uint16_t* knn, int k, int sx, int ry, int rz, uint16_t* brick_nodes, uint8_t* brick_count,
                                     uint32_t* brick_pairs) {
    const int nbx = (sx + BRICK_X - 1) / BRICK_X, nby = (ry + BRICK_Y - 1) / BRICK_Y, nbz = (rz + BRICK_Z - 1) / BRICK_Z;
    for (int bx = 0; bx < nbx; ++bx)
        for (int by = 0; by < nby; ++by)
            for (int bz = 0; bz < nbz; ++bz) {
                const size_t b = ((size_t)bx * nby + by) * nbz + bz;
                std::vector<uint16_t> set;
                for (int x = bx * BRICK_X; x < std::min(sx, (bx + 1) * BRICK_X); ++x)
                    for (int y = by * BRICK_Y; y < std::min(ry, (by + 1) * BRICK_Y); ++y)
                        for (int z = bz * BRICK_Z; z < std::min(rz, (bz + 1) * BRICK_Z); ++z)
                            for (int j = 0; j < k; ++j) {
                                const uint16_t id = knn[(((size_t)x * ry + y) * rz + z) * k + j];
                                bool found = false;
                                for (uint16_t s : set) found |= (s == id);
                                if (!found) set.push_back(id);
                            }
                brick_count[b] = set.size() > (size_t)BRICK_MAXC ? 255 : (uint8_t)set.size();
                for (size_t t = 0; t < set.size() && t < (size_t)BRICK_MAXC; ++t) brick_nodes[b * BRICK_MAXC + t] = set[t];
                for (int t = 0; t < BRICK_PAIR_WORDS; ++t) brick_pairs[b * BRICK_PAIR_WORDS + t] = 0;
                if (set.size() > (size_t)BRICK_MAXC) continue;
                for (int x = bx * BRICK_X; x < std::min(sx, (bx + 1) * BRICK_X); ++x)
                    for (int y = by * BRICK_Y; y < std::min(ry, (by + 1) * BRICK_Y); ++y)
                        for (int z = bz * BRICK_Z; z < std::min(rz, (bz + 1) * BRICK_Z); ++z) {
                            int loc[DFB_MAX_K];
                            for (int j = 0; j < k; ++j) {
                                const uint16_t id = knn[(((size_t)x * ry + y) * rz + z) * k + j];
                                loc[j] = (int)(std::find(set.begin(), set.end(), id) - set.begin());
                            }
                            for (int a = 0; a < k; ++a)
                                for (int c2 = 0; c2 <= a; ++c2) {
                                    const int hi = std::max(loc[a], loc[c2]), lo = std::min(loc[a], loc[c2]);
                                    const int p = hi * (hi + 1) / 2 + lo;
                                    brick_pairs[b * BRICK_PAIR_WORDS + (p >> 5)] |= 1u << (p & 31);
                                }
                        }
            }
}

// per-voxel brick class for the tests: 0xFF mixed, else clamp mask; frus bits in brick_frus
static const uint16_t* g_brick_nodes = nullptr;
static const uint8_t* g_brick_count = nullptr;
static const uint32_t* g_brick_pairs = nullptr;
static uint8_t* g_brick_cls_vox = nullptr;
static int g_use_regions = 0;
static float g_region_dmax = 0.f;   // largest deviation bound among valid regions of the last call (diagnostic)
static float g_region_valid = 0.f;  // fraction of valid regions
static int g_region_resolved = 0;   // regions classified as a whole in the last call
static float* g_dev_log = nullptr;
static long g_dev_n = 0;
extern "C" void hs_set_dev_log(float* p) { g_dev_log = p; g_dev_n = 0; }
extern "C" long hs_dev_log_n() { return g_dev_n; }
static int g_use_quads = 1;
static long g_quad_settled = 0, g_quad_band = 0, g_quad_open = 0;
extern "C" void hs_set_quads(int on) { g_use_quads = on; g_quad_settled = g_quad_band = g_quad_open = 0; }
extern "C" void hs_quad_stats(long* out) { out[0] = g_quad_settled; out[1] = g_quad_band; out[2] = g_quad_open; }
extern "C" void hs_set_bricks(const uint16_t* nodes, const uint8_t* count, const uint32_t* pairs, uint8_t* cls_vox, int use_regions) {
    g_brick_nodes = nodes; g_brick_count = count; g_brick_pairs = pairs; g_brick_cls_vox = cls_vox; g_use_regions = use_regions;
}
extern "C" float hs_region_dmax() { return g_region_dmax; }
extern "C" float hs_region_valid() { return g_region_valid; }
extern "C" int hs_region_resolved() { return g_region_resolved; }

// host mirror of region_build_kernel + region_bounds_kernel (tsdf.cu)
static std::vector<float> build_region_records(const ProjParams& P) {
    const int sx = P.x1 - P.x0;
    const int nrx = (sx + REGION_X - 1) / REGION_X, nry = (P.ry + REGION_Y - 1) / REGION_Y, nrz = (P.rz + REGION_Z - 1) / REGION_Z;
    std::vector<float> rec((size_t)nrx * nry * nrz * REGION_REC_FLOATS, 0.f);
    g_region_dmax = 0.f;
    g_region_resolved = 0;
    int nvalid = 0;
    for (int rx = 0; rx < nrx; ++rx)
        for (int ry_ = 0; ry_ < nry; ++ry_)
            for (int rz_ = 0; rz_ < nrz; ++rz_) {
                float* out = rec.data() + (((size_t)rx * nry + ry_) * nrz + rz_) * REGION_REC_FLOATS;
                const int xlo = rx * REGION_X, ylo = ry_ * REGION_Y, zlo = rz_ * REGION_Z;
                const int xhi = std::min(xlo + REGION_X, sx) - 1, yhi = std::min(ylo + REGION_Y, P.ry) - 1, zhi = std::min(zlo + REGION_Z, P.rz) - 1;
                std::vector<int> nodes;
                std::vector<std::pair<int, int>> pairs;
                for (int x = xlo; x <= xhi; ++x)
                    for (int y = ylo; y <= yhi; ++y)
                        for (int z = zlo; z <= zhi; ++z) {
                            const size_t i = ((size_t)x * P.ry + y) * P.rz + z;
                            int loc[DFB_MAX_K];
                            for (int j = 0; j < P.k; ++j) {
                                const int id = P.knn[i * P.k + j];
                                auto it = std::find(nodes.begin(), nodes.end(), id);
                                if (it == nodes.end()) { nodes.push_back(id); loc[j] = (int)nodes.size() - 1; }
                                else loc[j] = (int)(it - nodes.begin());
                            }
                            for (int a = 0; a < P.k; ++a)
                                for (int c2 = 0; c2 <= a; ++c2) {
                                    const std::pair<int, int> pr(std::max(loc[a], loc[c2]), std::min(loc[a], loc[c2]));
                                    if (std::find(pairs.begin(), pairs.end(), pr) == pairs.end()) pairs.push_back(pr);
                                }
                        }
                if (nodes.empty() || nodes.size() > (size_t)REGION_MAXC) continue;
                const float c[3] = {0.5f * (xlo + xhi) + (float)P.x0, 0.5f * (ylo + yhi), 0.5f * (zlo + zhi)};
                const float h[3] = {0.5f * (xhi - xlo), 0.5f * (yhi - ylo), 0.5f * (zhi - zlo)};
                bool bad = false;
                std::vector<std::array<float, 8>> q(nodes.size());
                for (size_t t = 0; t < nodes.size(); ++t) {
                    const float4 r0 = P.node_rec[3 * (size_t)nodes[t]], r1 = P.node_rec[3 * (size_t)nodes[t] + 1], r2 = P.node_rec[3 * (size_t)nodes[t] + 2];
                    q[t] = {r1.x, r1.y, r1.z, r1.w, r2.x, r2.y, r2.z, r2.w};
                    const float ddx = fabsf(c[0] - r0.x) + h[0], ddy = fabsf(c[1] - r0.y) + h[1], ddz = fabsf(c[2] - r0.z) + h[2];
                    if (!((ddx * ddx + ddy * ddy + ddz * ddz) * r0.w > -125.f)) bad = true;
                }
                float Pref[12];
                if (!region_reference_map(&q[0][0], (int)nodes.size(), Pref)) bad = true;
                float dev[3] = {0.f, 0.f, 0.f};
                for (auto& pr : pairs)
                    if (!region_pair_bound(q[pr.first].data(), q[pr.second].data(), pr.first == pr.second, Pref, c, h, dev)) bad = true;
                for (int t = 0; t < 12; ++t) out[t] = Pref[t];
                for (int r = 0; r < 3; ++r) out[12 + r] = dev[r];
                BrickClass rc = brick_class_all_mixed(P.n_views);
                if (!bad) {
                    Box3 bx;
                    region_box(out, c, h, P.coord_mag, bx);
                    rc = box_classify_views(P, bx, REGION_MAX_RECT, SerialCtx());
                }
                out[15] = region_code(!bad, rc);
                if (!bad && !rc.mixed) ++g_region_resolved;
                if (!bad && rc.mixed && g_dev_log) { g_dev_log[g_dev_n++] = dev[0]; g_dev_log[g_dev_n++] = dev[1]; g_dev_log[g_dev_n++] = dev[2]; }
                if (!bad) { ++nvalid; g_region_dmax = std::max(g_region_dmax, std::max(dev[0], std::max(dev[1], dev[2]))); }
            }
    g_region_valid = (float)nvalid / (float)((size_t)nrx * nry * nrz);
    return rec;
}

template <int KMAX>
static void run_proj(ProjParams& P, int mode, uint8_t* cls_out) {
    const size_t plane = (size_t)P.ry * P.rz;
    uint32_t n_unc = 0;
    const bool bricks = mode == DFB_MODE_HYBRID && (P.rigid || (g_brick_nodes && g_brick_count)) && g_brick_cls_vox;
    const int nby = (P.ry + BRICK_Y - 1) / BRICK_Y, nbz = (P.rz + BRICK_Z - 1) / BRICK_Z;
    std::vector<BrickClass> brick_cache((size_t)((P.x1 - P.x0 + BRICK_X - 1) / BRICK_X) * nby * nbz, BrickClass{-1, 0, 0, 0});
    std::vector<float> region_rec;
    if (bricks && !P.rigid && g_use_regions) region_rec = build_region_records(P);
    const float* rrec = region_rec.empty() ? nullptr : region_rec.data();
    for (int xs = 0; xs < P.x1 - P.x0; ++xs)
        for (int y = 0; y < P.ry; ++y)
            for (int z = 0; z < P.rz; ++z) {
                const size_t i = xs * plane + (size_t)y * P.rz + z;
                int views = 0xff, m0 = 0, f0 = 0;   // what the brick's box test hands to the per-voxel tier
                if (bricks) {
                    const size_t bid = ((size_t)(xs / BRICK_X) * nby + y / BRICK_Y) * nbz + z / BRICK_Z;
                    if (brick_cache[bid].cls < 0)
                        brick_cache[bid] = brick_classify(P, g_brick_nodes, g_brick_count, g_brick_pairs, rrec, nby, nbz, xs / BRICK_X, y / BRICK_Y, z / BRICK_Z, SerialCtx());
                    const BrickClass& B = brick_cache[bid];
                    const int bc = B.cls, fr = B.frus;
                    g_brick_cls_vox[i] = (uint8_t)bc;
                    if (!B.mixed) {
                        if (bc) {
                            float v = P.tsdf[i], w = P.weight[i];
                            for (int vi = 0; vi < P.n_views; ++vi)
                                if (bc & (1 << vi)) clamp_update(v, w, P.tdist_f, P.wmax_f, (float)P.scale);
                            P.tsdf[i] = v; P.weight[i] = w;
                        }
                        if (P.mask_out) P.mask_out[i] = (uint8_t)bc;
                        if (P.frustum_out) P.frustum_out[i] = (uint8_t)fr;
                        if (cls_out) cls_out[i] = bc ? CLS_CLAMP : CLS_SKIP;
                        continue;
                    }
                    views = B.mixed; m0 = B.clamp; f0 = B.frus;
                }
                uint16_t ids[KMAX] = {0};
                for (int j = 0; j < P.k; ++j) ids[j] = P.knn[i * P.k + j];
                int m = 0, f = 0, cls = CLS_UNCERTAIN;
                bool pretested = false;
                if (bricks && rrec && g_use_quads && mode == DFB_MODE_HYBRID) {
                    // the quad pre-test of the update kernel (dfb_brick.h quad_pretest), on the quad this voxel belongs to
                    const int nry = (P.ry + REGION_Y - 1) / REGION_Y, nrz = (P.rz + REGION_Z - 1) / REGION_Z;
                    const float* rr = rrec + (((size_t)(xs / REGION_X) * nry + y / REGION_Y) * nrz + z / REGION_Z) * REGION_REC_FLOATS;
                    if (rr[15] > 0.5f) {
                        const int zq = z & ~3, nz = std::min(4, P.rz - zq);
                        QuadView qv[DFB_MAX_VIEWS];
                        for (int v = 0; v < P.n_views; ++v) quad_view_setup(P, rr, v, qv[v]);
                        int st[4], mq[4], fq[4];
                        quad_pretest(P, qv, xs + P.x0, y, zq, nz, views, m0, f0, st, mq, fq);
                        const int q = z - zq;
                        if (st[q] == QV_SKIP) { m = mq[q]; f = fq[q]; cls = m ? CLS_CLAMP : CLS_SKIP; pretested = true; ++g_quad_settled; }
                        else if (st[q] == QV_BAND) { cls = CLS_UNCERTAIN; m = f = 0; pretested = true; ++g_quad_band; }
                        else ++g_quad_open;
                    }
                }
                if (mode == DFB_MODE_HYBRID && !pretested) cls = voxel_projective_classify<KMAX>(P, xs + P.x0, y, z, ids, &m, &f, views, m0, f0);
                if (cls_out) cls_out[i] = (uint8_t)cls;
                float v = P.tsdf[i], w = P.weight[i];
                if (cls == CLS_UNCERTAIN) {
                    ++n_unc;
                    voxel_projective_exact(P, xs + P.x0, y, z, ids, &v, &w, &m, &f);
                } else if (m) {
                    for (int vi = 0; vi < P.n_views; ++vi)
                        if (m & (1 << vi)) clamp_update(v, w, P.tdist_f, P.wmax_f, (float)P.scale);
                }
                if (m) { P.tsdf[i] = v; P.weight[i] = w; }
                if (P.mask_out) P.mask_out[i] = (uint8_t)m;
                if (P.frustum_out) P.frustum_out[i] = (uint8_t)f;
            }
    if (P.counters) P.counters[0] = n_unc;
}

extern "C" int hs_tsdf_update_projective(const dfb_volume* vol, const dfb_warpfield* wf, const dfb_views* views,
                                         double tdist, double wmax, int mode, const dfb_workspace* ws,
                                         uint8_t* mask_out, uint8_t* frustum_out, uint8_t* cls_out) {
    ProjParams P;
    if (int r = build_projective(P, vol, wf, views, tdist, wmax, mode, ws, mask_out, frustum_out)) return r;
    if (P.k <= 4) run_proj<4>(P, mode, cls_out);
    else run_proj<8>(P, mode, cls_out);
    return DFB_OK;
}

extern "C" int hs_fuse_depth_rigid(const dfb_volume* vol, int tsdf_res, const float* depth, int rows, int cols,
                                   const double* lw34, const double* K, const double* Kinv, double scale,
                                   const double* center, double tdist, double wmax, int mode, const dfb_workspace* ws,
                                   uint8_t* mask_out, uint8_t* frustum_out, uint8_t* cls_out) {
    ProjParams P;
    if (int r = build_rigid(P, vol, tsdf_res, depth, rows, cols, lw34, K, Kinv, scale, center, tdist, wmax, mode, ws,
                            mask_out, frustum_out))
        return r;
    run_proj<4>(P, mode, cls_out);
    return DFB_OK;
}

template <int KMAX>
static void run_vol(VolParams& P, int mode, uint8_t* cls_out) {
    const size_t plane = (size_t)P.ry * P.rz;
    uint32_t n_unc = 0;
    for (int xs = 0; xs < P.x1 - P.x0; ++xs)
        for (int y = 0; y < P.ry; ++y)
            for (int z = 0; z < P.rz; ++z) {
                const size_t i = xs * plane + (size_t)y * P.rz + z;
                uint16_t ids[KMAX] = {0};
                for (int j = 0; j < P.k; ++j) ids[j] = P.knn[i * P.k + j];
                float wi = 0.f;
                int cls = CLS_UNCERTAIN;
                if (mode == DFB_MODE_HYBRID) cls = voxel_volume_classify<KMAX>(P, xs + P.x0, y, z, ids, &wi);
                if (cls_out) cls_out[i] = (uint8_t)cls;
                float v = P.tsdf[i], w = P.weight[i];
                bool upd = false;
                if (cls == CLS_UNCERTAIN) {
                    ++n_unc;
                    upd = voxel_volume_exact(P, xs + P.x0, y, z, ids, &v, &w);
                } else if (cls == CLS_CLAMP) {
                    upd = true;
                    if (P.k > 0) {
                        const float wt = (w == 0.f) ? wi : w;
                        v = (v * wt + fmul(P.tdist_f, wi)) / (wi + wt);
                        w = fminf(wi + wt, P.wmax_f);
                    } else {
                        v = (v * w + P.tdist_f) / (1.0f + w);
                        w = fminf(1.0f + w, P.wmax_f);
                    }
                }
                if (upd) { P.tsdf[i] = v; P.weight[i] = w; }
                if (P.mask_out) P.mask_out[i] = upd ? 1 : 0;
            }
    if (P.counters) P.counters[0] = n_unc;
}

extern "C" int hs_tsdf_update_volume(const dfb_volume* vol, const dfb_warpfield* wf, const float* curr, int cx, int cy,
                                     int cz, double tdist, double wmax, int mode, const dfb_workspace* ws,
                                     uint8_t* mask_out, uint8_t* cls_out) {
    VolParams P;
    if (int r = build_volume(P, vol, wf, curr, cx, cy, cz, tdist, wmax, mode, ws, mask_out)) return r;
    if (P.k <= 4) run_vol<4>(P, mode, cls_out);
    else run_vol<8>(P, mode, cls_out);
    return DFB_OK;
}

extern "C" int hs_warp_points(const float* pts, const float* normals, int64_t m, const int32_t* idx,
                              const dfb_warpfield* wf, double* out_pts, double* out_normals) {
    if (int r = validate_warpfield(wf, false)) return r;
    for (int64_t t = 0; t < m; ++t) {
        int ids[DFB_MAX_K];
        for (int j = 0; j < wf->k; ++j) ids[j] = idx[t * wf->k + j];
        double on[3];
        warp_ref(pts + 3 * t, normals ? normals + 3 * t : nullptr, ids, wf->k, wf->node_pos, wf->node_dq, wf->node_w,
                 wf->lw, wf->has_lw != 0, wf->lw_is_f32 != 0, out_pts + 3 * t, on, nullptr);
        if (normals && out_normals)
            for (int c = 0; c < 3; ++c) out_normals[3 * t + c] = on[c];
    }
    return DFB_OK;
}

// ---- Gauss-Newton path: the kernels' per-residual functions (dfb_gn.h), dense assembly for small problems -------
#include "../../dynamicfusion_body_b200/csrc/dfb_gn.h"

static GNParams hs_params(const dfb_gn_problem* p) {
    GNParams P;
    P.n_vert = p->n_vert; P.vertices = p->vertices; P.normals = p->normals; P.corr = p->corr; P.vert_knn = p->vert_knn;
    P.order = nullptr;
    P.n_nodes = p->n_nodes; P.k = p->k; P.node_pos = p->node_pos; P.node_w = p->node_w; P.node_nbr = p->node_nbr;
    for (int i = 0; i < 8; ++i) P.lw[i] = p->lw[i];
    P.lw_is_f32 = p->lw_is_f32;
    dq_to_affine(p->lw, P.A);
    P.rw = p->rw; P.huber = p->huber; P.f_scale = p->f_scale > 0 ? p->f_scale : 1.0;
    return P;
}

extern "C" int hs_gn_residuals(const dfb_gn_problem* prob, const double* x, int x_is_f32, double* f) {
    const GNParams P = hs_params(prob);
    for (int64_t i = 0; i < P.n_vert; ++i) f[i] = data_residual_ref(P, x, x_is_f32 != 0, P.lw, P.lw_is_f32 != 0, i);
    for (int i = 0; i < P.n_nodes; ++i)
        for (int jj = 0; jj < P.k; ++jj) reg_residual_ref(P, x, x_is_f32 != 0, i, jj, f + P.n_vert + 3 * ((int64_t)i * P.k + jj));
    return 0;
}

extern "C" int hs_gn_residuals_lw(const dfb_gn_problem* prob, const double* dq, int dq_is_f32, const double* lw, int lw_is_f32, double* f) {
    const GNParams P = hs_params(prob);
    for (int64_t i = 0; i < P.n_vert; ++i) f[i] = data_residual_ref(P, dq, dq_is_f32 != 0, lw, lw_is_f32 != 0, i);
    return 0;
}

// dense H [8N][8N], g [8N], cost[2] with the same per-residual functions / weights as normal_eq_*_kernel
extern "C" int hs_gn_normal_eq_dense(const dfb_gn_problem* prob, const double* x, double* H, double* g, double* cost) {
    const GNParams P = hs_params(prob);
    const int64_t n8 = 8 * (int64_t)P.n_nodes;
    for (int64_t i = 0; i < n8 * n8; ++i) H[i] = 0;
    for (int64_t i = 0; i < n8; ++i) g[i] = 0;
    cost[0] = cost[1] = 0;
    for (int64_t i = 0; i < P.n_vert; ++i) {
        double r, gv[8], wts[DFB_MAX_K];
        data_residual_jac(P, x, i, &r, gv, wts);
        const double om = huber_weight(r, P.huber, P.f_scale);
        const int32_t* ids = P.vert_knn + i * P.k;
        for (int a = 0; a < P.k; ++a) {
            for (int b = 0; b < P.k; ++b)
                for (int rr = 0; rr < 8; ++rr)
                    for (int cc = 0; cc < 8; ++cc) H[(8 * (int64_t)ids[a] + rr) * n8 + 8 * ids[b] + cc] += om * wts[a] * wts[b] * gv[rr] * gv[cc];
            for (int c = 0; c < 8; ++c) g[8 * (int64_t)ids[a] + c] += om * wts[a] * r * gv[c];
        }
        cost[0] += huber_rho(r, P.huber, P.f_scale);
        cost[1] += 0.5 * r * r;
    }
    for (int i = 0; i < P.n_nodes; ++i)
        for (int jj = 0; jj < P.k; ++jj) {
            double r[3], Ji[3][8], Jj[3][8];
            const int j = reg_residual_jac(P, x, i, jj, r, Ji, Jj);
            for (int t = 0; t < 3; ++t) {
                const double om = huber_weight(r[t], P.huber, P.f_scale);
                cost[0] += huber_rho(r[t], P.huber, P.f_scale);
                cost[1] += 0.5 * r[t] * r[t];
                if (j == i) continue;
                for (int a = 0; a < 8; ++a) {
                    for (int b = 0; b < 8; ++b) {
                        H[(8 * (int64_t)i + a) * n8 + 8 * i + b] += om * Ji[t][a] * Ji[t][b];
                        H[(8 * (int64_t)i + a) * n8 + 8 * j + b] += om * Ji[t][a] * Jj[t][b];
                        H[(8 * (int64_t)j + a) * n8 + 8 * i + b] += om * Jj[t][a] * Ji[t][b];
                        H[(8 * (int64_t)j + a) * n8 + 8 * j + b] += om * Jj[t][a] * Jj[t][b];
                    }
                    g[8 * (int64_t)i + a] += om * Ji[t][a] * r[t];
                    g[8 * (int64_t)j + a] += om * Jj[t][a] * r[t];
                }
            }
        }
    return 0;
}

extern "C" int hs_gn_lw_normal_eq(const dfb_gn_problem* prob, const double* dq, const double* lw, double* H8, double* g8, double* cost) {
    const GNParams P = hs_params(prob);
    for (int i = 0; i < 64; ++i) H8[i] = 0;
    for (int i = 0; i < 8; ++i) g8[i] = 0;
    cost[0] = cost[1] = 0;
    for (int64_t i = 0; i < P.n_vert; ++i) {
        double r, J[8];
        lw_residual_jac(P, dq, lw, i, &r, J);
        const double om = huber_weight(r, P.huber, P.f_scale);
        for (int a = 0; a < 8; ++a) {
            for (int b = 0; b < 8; ++b) H8[a * 8 + b] += om * J[a] * J[b];
            g8[a] += om * J[a] * r;
        }
        cost[0] += huber_rho(r, P.huber, P.f_scale);
        cost[1] += 0.5 * r * r;
    }
    return 0;
}

// ---- surface extraction (csrc/dfb_mc.h): the loops of mc.cu's count / scan / emit kernels with HOST pointers ------------------
#include "../../dynamicfusion_body_b200/csrc/dfb_mc.h"

extern "C" float hs_mc_level(const float* vol, int64_t n) {
    float lo = INFINITY, hi = -INFINITY;
    for (int64_t i = 0; i < n; ++i) { lo = fminf(lo, vol[i]); hi = fmaxf(hi, vol[i]); }
    return (float)(0.5 * ((double)lo + (double)hi));
}

extern "C" void hs_mc_count(const float* vol, int rx, int ry, int rz, int step, int x_origin, float level, McChunk* chunks, int32_t* row_voff,
                            int32_t* row_toff) {
    McGrid g;
    mc_grid_init(g, vol, rx, ry, rz, step, level, x_origin);
    const int rows = g.nx * g.ny;
    for (int row = 0; row < rows; ++row) {
        const int i = row / g.ny, j = row - i * g.ny;
        const bool cell_row = i + 1 < g.nx && j + 1 < g.ny;
        int nv = 0, nt = 0;
        for (int c = 0; c < g.ncz; ++c) {
            McChunk rec;
            rec.voff = nv; rec.m[0] = rec.m[1] = rec.m[2] = 0;
            for (int lane = 0; lane < 32; ++lane) {
                const int k = 32 * c + lane;
                if (k >= g.nz) break;
                const uint32_t bits = mc_edge_bits(g, i, j, k, mc_val(g, i, j, k));
                for (int d = 0; d < 3; ++d) rec.m[d] |= ((bits >> d) & 1u) << lane;
                if (cell_row && k + 1 < g.nz) {
                    float v[8];
                    const int cs = mc_cell_case(g, i, j, k, v);
                    if (cs != 0 && cs != 255) nt += mc_cell_tris(cs, v, g.level, i + g.xs0, j, k, nullptr);
                }
            }
            chunks[(size_t)row * g.ncz + c] = rec;
            nv += mc_popc(rec.m[0]) + mc_popc(rec.m[1]) + mc_popc(rec.m[2]);
        }
        row_voff[row] = nv; row_toff[row] = nt;
    }
    int32_t a = 0, b = 0;
    for (int row = 0; row < rows; ++row) {
        const int32_t na = row_voff[row], nb = row_toff[row];
        row_voff[row] = a; row_toff[row] = b;
        a += na; b += nb;
    }
    row_voff[rows] = a; row_toff[rows] = b;
}

extern "C" void hs_mc_emit(const float* vol, int rx, int ry, int rz, int step, int x_origin, float level, const McChunk* chunks,
                           const int32_t* row_voff, const int32_t* row_toff, float* verts, float* normals, float* values, int32_t* faces) {
    McGrid g;
    mc_grid_init(g, vol, rx, ry, rz, step, level, x_origin);
    const int rows = g.nx * g.ny;
    for (int row = 0; row < rows; ++row) {
        const int i = row / g.ny, j = row - i * g.ny;
        const bool cell_row = i + 1 < g.nx && j + 1 < g.ny;
        int tbase = row_toff[row];
        for (int c = 0; c < g.ncz; ++c) {
            const McChunk rec = chunks[(size_t)row * g.ncz + c];
            for (int lane = 0; lane < 32; ++lane) {
                const int k = 32 * c + lane;
                if (k >= g.nz) break;
                if (((rec.m[0] | rec.m[1] | rec.m[2]) >> lane) & 1u) {
                    int id = mc_vertex_id(g, chunks, row_voff, i, j, k, 0);
                    for (int d = 0; d < 3; ++d)
                        if ((rec.m[d] >> lane) & 1u) {
                            float val;
                            mc_vertex(g, i, j, k, d, verts + 3 * (size_t)id, normals + 3 * (size_t)id, val);
                            values[id] = val;
                            ++id;
                        }
                }
                if (cell_row && k + 1 < g.nz) {
                    float v[8];
                    int8_t edges[3 * DFB_MC_MAX_TRIS];
                    const int cs = mc_cell_case(g, i, j, k, v);
                    const int nt = (cs != 0 && cs != 255) ? mc_cell_tris(cs, v, g.level, i + g.xs0, j, k, edges) : 0;
                    for (int t = 0; t < nt; ++t)
                        for (int q = 0; q < 3; ++q)
                            faces[3 * (size_t)(tbase + t) + q] = mc_cell_edge_vertex(g, chunks, row_voff, i, j, k, edges[3 * t + q]);
                    tbase += nt;
                }
            }
        }
    }
}
