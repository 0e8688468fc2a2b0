// dfb_gn.h -- per-residual arithmetic of the warp-field least-squares path, shared by the
// CUDA kernels (gn.cu) and the CPU logic tests (tests/hostshim/).
//
//   residual VALUES  : reference arithmetic (dfb_math.h exact tier) -- Fusion.computef / computef_lw
//   residual JACOBIANS: analytic, float64, on the smooth model (the reference's float32 roundings Q3 are not
//                       differentiated).  The reference never forms them: scipy finite-differences computef
//                       (core/fusion.py:382-392), which costs 93.9 % of its solve time.
#pragma once
#include "dfb_math.h"

namespace dfb {

struct GNParams {
    int64_t n_vert;
    const float* vertices;   // [V][3]  Fusion._vertices (float32, marching-cubes output in the reference)
    const float* normals;    // [V][3]  Fusion._normals
    const double* corr;      // [V][3]  Fusion._correspondences
    const int32_t* vert_knn; // [V][k]  Fusion._neighbor_look_up
    const int32_t* order;    // [V] or null: order in which the assembly walks the data residuals (sorted by node tuple)
    int n_nodes, k;
    const float* node_pos;   // [N][3]
    const float* node_w;     // [N]
    const int32_t* node_nbr; // [N][k]  _neighbor_look_up[_nodes[i][0]]  (core/fusion.py:477)
    double lw[8];
    int lw_is_f32;
    double A[12];            // affine of lw (smooth model)
    double rw;               // regularization_weight
    int huber;               // IRLS weights min(1, f_scale/|f|)
    double f_scale;
};

// W(q,p) = dqb_warp closed form for a (possibly non-unit) dq
DFB_HD void W_apply(const double* q, const double* p, bool rot_only, double* o) {
    const double w = q[0], x = q[1], y = q[2], z = q[3];
    const double dw = rot_only ? 0.0 : q[4], dx = rot_only ? 0.0 : q[5], dy = rot_only ? 0.0 : q[6], dz = rot_only ? 0.0 : q[7];
    const double s = w * w - (x * x + y * y + z * z);
    const double vp = x * p[0] + y * p[1] + z * p[2];
    const double cx = y * p[2] - z * p[1], cy = z * p[0] - x * p[2], cz = x * p[1] - y * p[0];
    o[0] = s * p[0] + 2.0 * (vp * x + w * cx + (w * dx - dw * x + (y * dz - z * dy)));
    o[1] = s * p[1] + 2.0 * (vp * y + w * cy + (w * dy - dw * y + (z * dx - x * dz)));
    o[2] = s * p[2] + 2.0 * (vp * z + w * cz + (w * dz - dw * z + (x * dy - y * dx)));
}

// J[i][c] = d W_i / d q_c
DFB_HD void dW_dq(const double* q, const double* p, bool rot_only, double J[3][8]) {
    const double w = q[0];
    const double v[3] = {q[1], q[2], q[3]};
    const double dw = rot_only ? 0.0 : q[4];
    const double dv[3] = {rot_only ? 0.0 : q[5], rot_only ? 0.0 : q[6], rot_only ? 0.0 : q[7]};
    const double vxp[3] = {v[1] * p[2] - v[2] * p[1], v[2] * p[0] - v[0] * p[2], v[0] * p[1] - v[1] * p[0]};
    const double vp = v[0] * p[0] + v[1] * p[1] + v[2] * p[2];
    // skew matrices: [a]x[i][j]
    const double Px[3][3] = {{0, -p[2], p[1]}, {p[2], 0, -p[0]}, {-p[1], p[0], 0}};
    const double Dx[3][3] = {{0, -dv[2], dv[1]}, {dv[2], 0, -dv[0]}, {-dv[1], dv[0], 0}};
    const double Vx[3][3] = {{0, -v[2], v[1]}, {v[2], 0, -v[0]}, {-v[1], v[0], 0}};
    for (int i = 0; i < 3; ++i) {
        J[i][0] = 2.0 * w * p[i] + 2.0 * vxp[i] + 2.0 * dv[i];
        for (int j = 0; j < 3; ++j) {
            const double id = (i == j) ? 1.0 : 0.0;
            J[i][1 + j] = -2.0 * p[i] * v[j] + 2.0 * v[i] * p[j] + 2.0 * vp * id - 2.0 * w * Px[i][j] - 2.0 * dw * id - 2.0 * Dx[i][j];
            J[i][5 + j] = rot_only ? 0.0 : (2.0 * w * id + 2.0 * Vx[i][j]);
        }
        J[i][4] = rot_only ? 0.0 : -2.0 * v[i];
    }
}

DFB_HD double huber_weight(double f, int huber, double f_scale) {
    if (!huber) return 1.0;
    const double a = fabs(f) / f_scale;
    return a <= 1.0 ? 1.0 : 1.0 / a;
}

DFB_HD double huber_rho(double f, int huber, double f_scale) {  // 0.5 * f_scale^2 * rho((f/f_scale)^2)
    if (!huber) return 0.5 * f * f;
    const double z = (f / f_scale) * (f / f_scale);
    return 0.5 * f_scale * f_scale * (z <= 1.0 ? z : 2.0 * sqrt(z) - 1.0);
}

// Smooth-model data residual of vertex i: r, g = d r / d b (8), wts (k) so that d r / d dq_a = wts[a] * g.
// x: node dual quaternions [N][8] float64.
DFB_HDN void data_residual_jac(const GNParams& P, const double* x, int64_t i, double* r_out, double* g, double* wts) {
    const double p[3] = {(double)P.vertices[3 * i], (double)P.vertices[3 * i + 1], (double)P.vertices[3 * i + 2]};
    const double n[3] = {(double)P.normals[3 * i], (double)P.normals[3 * i + 1], (double)P.normals[3 * i + 2]};
    double b[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int a = 0; a < P.k; ++a) {
        const int id = P.vert_knn[i * P.k + a];
        const double dx = p[0] - (double)P.node_pos[3 * id], dy = p[1] - (double)P.node_pos[3 * id + 1], dz = p[2] - (double)P.node_pos[3 * id + 2];
        const double w2 = 2.0 * (double)P.node_w[id];
        const double w = exp(-(dx * dx + dy * dy + dz * dz) / (w2 * w2));
        wts[a] = w;
        for (int c = 0; c < 8; ++c) b[c] += w * x[8 * (size_t)id + c];
    }
    double s2 = 0;
    for (int c = 0; c < 8; ++c) s2 += b[c] * b[c];
    const double s = sqrt(s2);
    double bh[8];
    for (int c = 0; c < 8; ++c) bh[c] = b[c] / s;
    double v1[3], n1[3], v2[3], n2[3], e[3];
    W_apply(bh, p, false, v1);
    W_apply(bh, n, true, n1);
    for (int r = 0; r < 3; ++r) {
        v2[r] = P.A[4 * r] * v1[0] + P.A[4 * r + 1] * v1[1] + P.A[4 * r + 2] * v1[2] + P.A[4 * r + 3];
        n2[r] = P.A[4 * r] * n1[0] + P.A[4 * r + 1] * n1[1] + P.A[4 * r + 2] * n1[2];
        e[r] = v2[r] - P.corr[3 * i + r];
    }
    *r_out = n2[0] * e[0] + n2[1] * e[1] + n2[2] * e[2];
    double an[3], ae[3];  // A^T n2, A^T e
    for (int c = 0; c < 3; ++c) {
        an[c] = P.A[c] * n2[0] + P.A[4 + c] * n2[1] + P.A[8 + c] * n2[2];
        ae[c] = P.A[c] * e[0] + P.A[4 + c] * e[1] + P.A[8 + c] * e[2];
    }
    double Jv[3][8], Jn[3][8], gt[8];
    dW_dq(bh, p, false, Jv);
    dW_dq(bh, n, true, Jn);
    double bg = 0;
    for (int c = 0; c < 8; ++c) {
        gt[c] = an[0] * Jv[0][c] + an[1] * Jv[1][c] + an[2] * Jv[2][c] + ae[0] * Jn[0][c] + ae[1] * Jn[1][c] + ae[2] * Jn[2][c];
        bg += bh[c] * gt[c];
    }
    for (int c = 0; c < 8; ++c) g[c] = (gt[c] - bh[c] * bg) / s;
}

// Smooth-model regularisation residual (node i, neighbour slot jj): r[3], Ji[3][8] (w.r.t. dq_i), Jj (w.r.t. dq_j).
DFB_HDN int reg_residual_jac(const GNParams& P, const double* x, int i, int jj, double* r, double Ji[3][8], double Jj[3][8]) {
    const int j = P.node_nbr[i * P.k + jj];
    const double pj[3] = {(double)P.node_pos[3 * j], (double)P.node_pos[3 * j + 1], (double)P.node_pos[3 * j + 2]};
    const double wi = (double)P.node_w[i], wj = (double)P.node_w[j];
    const double c = P.rw * (wi > wj ? wi : wj);
    double a[3], b[3];
    W_apply(x + 8 * (size_t)i, pj, false, a);
    W_apply(x + 8 * (size_t)j, pj, false, b);
    dW_dq(x + 8 * (size_t)i, pj, false, Ji);
    dW_dq(x + 8 * (size_t)j, pj, false, Jj);
    for (int q = 0; q < 3; ++q) {
        r[q] = c * (a[q] - b[q]);
        for (int t = 0; t < 8; ++t) { Ji[q][t] *= c; Jj[q][t] *= -c; }
    }
    return j;
}

// Smooth-model data residual as a function of the global rigid dq lw: r and J (8).  dq: node transforms [N][8] f64.
DFB_HDN void lw_residual_jac(const GNParams& P, const double* dq, const double* lw, int64_t i, double* r_out, double* J) {
    const double p[3] = {(double)P.vertices[3 * i], (double)P.vertices[3 * i + 1], (double)P.vertices[3 * i + 2]};
    const double n[3] = {(double)P.normals[3 * i], (double)P.normals[3 * i + 1], (double)P.normals[3 * i + 2]};
    double b[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int a = 0; a < P.k; ++a) {
        const int id = P.vert_knn[i * P.k + a];
        const double dx = p[0] - (double)P.node_pos[3 * id], dy = p[1] - (double)P.node_pos[3 * id + 1], dz = p[2] - (double)P.node_pos[3 * id + 2];
        const double w2 = 2.0 * (double)P.node_w[id];
        const double w = exp(-(dx * dx + dy * dy + dz * dz) / (w2 * w2));
        for (int c = 0; c < 8; ++c) b[c] += w * dq[8 * (size_t)id + c];
    }
    double s2 = 0;
    for (int c = 0; c < 8; ++c) s2 += b[c] * b[c];
    const double s = sqrt(s2);
    double bh[8];
    for (int c = 0; c < 8; ++c) bh[c] = (P.k > 0) ? b[c] / s : (c == 0 ? 1.0 : 0.0);
    double v1[3], n1[3], v2[3], n2[3], e[3];
    W_apply(bh, p, false, v1);
    W_apply(bh, n, true, n1);
    W_apply(lw, v1, false, v2);
    W_apply(lw, n1, true, n2);
    for (int r = 0; r < 3; ++r) e[r] = v2[r] - P.corr[3 * i + r];
    *r_out = n2[0] * e[0] + n2[1] * e[1] + n2[2] * e[2];
    double Jv[3][8], Jn[3][8];
    dW_dq(lw, v1, false, Jv);
    dW_dq(lw, n1, true, Jn);
    for (int c = 0; c < 8; ++c)
        J[c] = n2[0] * Jv[0][c] + n2[1] * Jv[1][c] + n2[2] * Jv[2][c] + e[0] * Jn[0][c] + e[1] * Jn[1][c] + e[2] * Jn[2][c];
}

// ---- reference-arithmetic residual values ------------------------------------------------------------------
// dq_blend_ref variant whose node transforms are float64 values (x); x_is_f32: they are float32 values in the
// reference (`w * dg_dq` then rounds in float32).
DFB_HDN void dq_blend_ref_x(const float* p, const int32_t* ids, int k, const float* node_pos, const double* x, bool x_is_f32,
                            const float* node_w, double* se3) {
    double b[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < k; ++i) {
        const int id = ids[i];
        const float* np_ = node_pos + 3 * (size_t)id;
        const float nrm = norm3_f32_ref(p[0], p[1], p[2], np_[0], np_[1], np_[2]);
        const float two_w = fmul(2.0f, node_w[id]);
        const float q = fdiv(nrm, two_w);
        const float arg = fmul(-1.0f, fmul(q, q));
        const double w = exp((double)arg);
        const double* dqi = x + 8 * (size_t)id;
        if (x_is_f32) {
            const float wf = (float)w;
            for (int c = 0; c < 8; ++c) b[c] = dadd(b[c], (double)fmul(wf, (float)dqi[c]));
        } else {
            for (int c = 0; c < 8; ++c) b[c] = dadd(b[c], dmul(w, dqi[c]));
        }
    }
    double s = 0.0;
    for (int c = 0; c < 8; ++c) s = dadd(s, dmul(b[c], b[c]));
    const double nrm8 = dsqrt(s);
    if (nrm8 == 0.0) {
        se3[0] = 1.0;
        for (int c = 1; c < 8; ++c) se3[c] = 0.0;
        return;
    }
    for (int c = 0; c < 8; ++c) se3[c] = ddiv(b[c], nrm8);
}

// n'.(v' - corr) with (v',n') = warp(vertex, dqs[loc], loc, normal, m_lw) (core/fusion.py:469-470)
DFB_HDN double data_residual_ref(const GNParams& P, const double* x, bool x_is_f32, const double* lw, bool lw_is_f32, int64_t i) {
    const float* p = P.vertices + 3 * i;
    double se3[8], v1[3], v2[3], n1[3], n2[3];
    const double pd[3] = {(double)p[0], (double)p[1], (double)p[2]};
    const double nd[3] = {(double)P.normals[3 * i], (double)P.normals[3 * i + 1], (double)P.normals[3 * i + 2]};
    dq_blend_ref_x(p, P.vert_knn + i * P.k, P.k, P.node_pos, x, x_is_f32, P.node_w, se3);
    dqb_warp_ref(se3, false, pd, v1);
    dqb_warp_ref(lw, lw_is_f32, v1, v2);
    dqb_warp_normal_ref(se3, nd, n1);
    dqb_warp_normal_ref(lw, n1, n2);
    // np.dot(n_warped, vert_warped - corr)
    const double e0 = dsub(v2[0], P.corr[3 * i]), e1 = dsub(v2[1], P.corr[3 * i + 1]), e2 = dsub(v2[2], P.corr[3 * i + 2]);
    return dadd(dadd(dmul(n2[0], e0), dmul(n2[1], e1)), dmul(n2[2], e2));
}

// rw * max(w_i,w_j) * (dqb_warp(dq_i, v_j) - dqb_warp(dq_j, v_j))  (core/fusion.py:480-482)
DFB_HDN void reg_residual_ref(const GNParams& P, const double* x, bool x_is_f32, int i, int jj, double* r) {
    const int j = P.node_nbr[i * P.k + jj];
    const double pj[3] = {(double)P.node_pos[3 * j], (double)P.node_pos[3 * j + 1], (double)P.node_pos[3 * j + 2]};
    double a[3], b[3];
    dqb_warp_ref(x + 8 * (size_t)i, x_is_f32, pj, a);
    dqb_warp_ref(x + 8 * (size_t)j, x_is_f32, pj, b);
    const double wi = (double)P.node_w[i], wj = (double)P.node_w[j];
    const double c = dmul(P.rw, wi > wj ? wi : wj);
    for (int q = 0; q < 3; ++q) r[q] = dmul(c, dsub(a[q], b[q]));
}

}  // namespace dfb
