// emia_ellipse.cuh — least-squares ellipse fit on contour vertices (host/device).
//
// Replaces (reference call site): cv2.fitEllipse(c) at src/utils/measurements.py:177, of which the reference
// uses only the two axis lengths ((x, y), (major_axis, minor_axis), angle = ellipse — width <= height, Q3).
//
// Method (OpenCV's general-conic two-stage fit): centre the points on their float32 mean, scale by
// 100 / sum(|dx|+|dy|); stage 1: least squares for  -A x^2 - B y^2 - C xy + D x + E y = 10000 (5 unknowns);
// the conic centre follows from the 2x2 system [2A C; C 2B] r = [D; E]; stage 2: least squares for
// A' (x-rx)^2 + B' (y-ry)^2 + C' (x-rx)(y-ry) = 1 (3 unknowns); axes from A', B', C'.
// OpenCV solves both systems with an SVD in fp64; here they are solved with a row-streaming Givens QR in fp64
// (no n x 5 matrix is ever stored), which is backward stable like the SVD solve; outputs are rounded to
// float32 exactly as OpenCV does.  Ill-conditioned input (sigma_max * FLT_EPSILON > sigma_min) takes OpenCV's
// deterministic perturbation path.
#pragma once
#include "emia_common.cuh"
#include "emia_contour.cuh"

struct EmiaEllipse { float cx, cy, w, h, angle; int ok; };

// Incorporate one row (a[0..K-1] | rhs) into upper-triangular R (K x K, row-major) and qtb with Givens rotations.
template <int K>
EMIA_HD void emia_givens_add_row(double* R, double* qtb, double* a, double rhs) {
    for (int k = 0; k < K; ++k) {
        const double x = a[k];
        if (x == 0.0) continue;
        const double r = R[k * K + k];
        const double h = hypot(r, x);
        const double c = r / h, s = x / h;
        R[k * K + k] = h;
        for (int j = k + 1; j < K; ++j) {
            const double rkj = R[k * K + j], aj = a[j];
            R[k * K + j] = c * rkj + s * aj;
            a[j] = -s * rkj + c * aj;
        }
        const double qb = qtb[k];
        qtb[k] = c * qb + s * rhs;
        rhs = -s * qb + c * rhs;
    }
}
template <int K>
EMIA_HD void emia_back_subst(const double* R, const double* qtb, double* x) {
    for (int k = K - 1; k >= 0; --k) {
        double v = qtb[k];
        for (int j = k + 1; j < K; ++j) v -= R[k * K + j] * x[j];
        x[k] = (R[k * K + k] != 0.0) ? v / R[k * K + k] : 0.0;
    }
}
// extreme singular values of upper-triangular R (K x K) by one-sided Jacobi on a copy
template <int K>
EMIA_HD void emia_sv_extremes(const double* R, double* smax, double* smin) {
    double M[K * K];
    for (int i = 0; i < K * K; ++i) M[i] = R[i];
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off = 0.0;
        for (int p = 0; p < K - 1; ++p)
            for (int q = p + 1; q < K; ++q) {
                double alpha = 0, beta = 0, gamma = 0;
                for (int i = 0; i < K; ++i) {
                    alpha += M[i * K + p] * M[i * K + p];
                    beta += M[i * K + q] * M[i * K + q];
                    gamma += M[i * K + p] * M[i * K + q];
                }
                if (gamma == 0.0) continue;
                off += fabs(gamma) / sqrt(alpha * beta + 1e-300);
                const double zeta = (beta - alpha) / (2.0 * gamma);
                const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                for (int i = 0; i < K; ++i) {
                    const double mp = M[i * K + p], mq = M[i * K + q];
                    M[i * K + p] = c * mp - s * mq;
                    M[i * K + q] = s * mp + c * mq;
                }
            }
        if (off < 1e-15) break;
    }
    double mx = 0.0, mn = 1e300;
    for (int j = 0; j < K; ++j) {
        double nrm = 0.0;
        for (int i = 0; i < K; ++i) nrm += M[i * K + j] * M[i * K + j];
        nrm = sqrt(nrm);
        if (nrm > mx) mx = nrm;
        if (nrm < mn) mn = nrm;
    }
    *smax = mx; *smin = mn;
}

EMIA_HD void emia_ellipse_ofs(int i, float eps, float* ox, float* oy) {
    *ox = (float)(((i & 1) * 2 - 1)) * eps;
    *oy = (float)(((i & 2) - 1)) * eps;
}

// ---- fast least squares: normal equations + Cholesky + iterative refinement ---------------------------------------
// Solves min |A x - rhs| for K unknowns from the accumulated Gram matrix G = A^T A (row-major, full) and g = A^T rhs.
// Returns det(G) / trace(G)^K, a cheap lower bound of lambda_min / lambda_max (0 when the factorisation breaks down).
template <int K>
EMIA_HD double emia_chol_solve(const double* G, const double* g, double* L, double* x) {
    double tr = 0.0;
    for (int i = 0; i < K; ++i) tr += G[i * K + i];
    double det = 1.0;
    for (int j = 0; j < K; ++j) {
        double d = G[j * K + j];
        for (int k = 0; k < j; ++k) d -= L[j * K + k] * L[j * K + k];
        if (!(d > 0.0)) return 0.0;
        const double ljj = sqrt(d);
        L[j * K + j] = ljj;
        det *= d;
        for (int i = j + 1; i < K; ++i) {
            double v = G[i * K + j];
            for (int k = 0; k < j; ++k) v -= L[i * K + k] * L[j * K + k];
            L[i * K + j] = v / ljj;
        }
    }
    // forward / backward substitution
    double y[K];
    for (int i = 0; i < K; ++i) {
        double v = g[i];
        for (int k = 0; k < i; ++k) v -= L[i * K + k] * y[k];
        y[i] = v / L[i * K + i];
    }
    for (int i = K - 1; i >= 0; --i) {
        double v = y[i];
        for (int k = i + 1; k < K; ++k) v -= L[k * K + i] * x[k];
        x[i] = v / L[i * K + i];
    }
    double trk = 1.0;
    for (int i = 0; i < K; ++i) trk *= tr;
    return det / trk;
}
template <int K>
EMIA_HD void emia_chol_resolve(const double* L, const double* g, double* x) {
    double y[K];
    for (int i = 0; i < K; ++i) {
        double v = g[i];
        for (int k = 0; k < i; ++k) v -= L[i * K + k] * y[k];
        y[i] = v / L[i * K + i];
    }
    for (int i = K - 1; i >= 0; --i) {
        double v = y[i];
        for (int k = i + 1; k < K; ++k) v -= L[k * K + i] * x[k];
        x[i] = v / L[i * K + i];
    }
}

struct EmiaEllipsePts {
    const uint32_t* pts; int n; float cxf, cyf; double scale; int perturbed; float eps;
};
EMIA_HD void emia_ellipse_pt(const EmiaEllipsePts& P, int i, double* px, double* py) {
    float fx = (float)EMIA_PT_X(P.pts[i]), fy = (float)EMIA_PT_Y(P.pts[i]);
    if (P.perturbed) { float ox, oy; emia_ellipse_ofs(i, P.eps, &ox, &oy); fx = fx + ox; fy = fy + oy; }
    const float dx = fx - P.cxf, dy = fy - P.cyf;
    *px = dx * P.scale; *py = dy * P.scale;
}

// lower bound on lambda_min/lambda_max below which the fast path hands over to the rotation-based (QR + Jacobi) path
#define EMIA_ELLIPSE_FAST_MIN_RATIO 1e-11

// stage 1, fast: returns 1 and fills gfp when the system is comfortably well conditioned
EMIA_HD int emia_ellipse_stage1_fast(const EmiaEllipsePts& P, double* gfp) {
    double G[25], g[5], L[25];
    for (int i = 0; i < 25; ++i) { G[i] = 0.0; L[i] = 0.0; }
    for (int i = 0; i < 5; ++i) g[i] = 0.0;
    for (int i = 0; i < P.n; ++i) {
        double px, py; emia_ellipse_pt(P, i, &px, &py);
        const double row[5] = {-px * px, -py * py, -px * py, px, py};
        for (int a = 0; a < 5; ++a) {
            for (int b = a; b < 5; ++b) G[a * 5 + b] += row[a] * row[b];
            g[a] += row[a] * 10000.0;
        }
    }
    for (int a = 0; a < 5; ++a) for (int b = 0; b < a; ++b) G[a * 5 + b] = G[b * 5 + a];
    const double ratio = emia_chol_solve<5>(G, g, L, gfp);
    if (!(ratio > EMIA_ELLIPSE_FAST_MIN_RATIO)) return 0;
    // two steps of iterative refinement with residuals taken against the rows themselves
    for (int it = 0; it < 2; ++it) {
        double r5[5] = {0, 0, 0, 0, 0};
        for (int i = 0; i < P.n; ++i) {
            double px, py; emia_ellipse_pt(P, i, &px, &py);
            const double row[5] = {-px * px, -py * py, -px * py, px, py};
            double res = 10000.0;
            for (int a = 0; a < 5; ++a) res -= row[a] * gfp[a];
            for (int a = 0; a < 5; ++a) r5[a] += row[a] * res;
        }
        double dx5[5];
        emia_chol_resolve<5>(L, r5, dx5);
        for (int a = 0; a < 5; ++a) gfp[a] += dx5[a];
    }
    return 1;
}
EMIA_HD int emia_ellipse_stage2_fast(const EmiaEllipsePts& P, const double* rp, double* gfp) {
    double G[9], g[3], L[9];
    for (int i = 0; i < 9; ++i) { G[i] = 0.0; L[i] = 0.0; }
    for (int i = 0; i < 3; ++i) g[i] = 0.0;
    for (int i = 0; i < P.n; ++i) {
        double px, py; emia_ellipse_pt(P, i, &px, &py);
        const double row[3] = {(px - rp[0]) * (px - rp[0]), (py - rp[1]) * (py - rp[1]), (px - rp[0]) * (py - rp[1])};
        for (int a = 0; a < 3; ++a) {
            for (int b = a; b < 3; ++b) G[a * 3 + b] += row[a] * row[b];
            g[a] += row[a];
        }
    }
    for (int a = 0; a < 3; ++a) for (int b = 0; b < a; ++b) G[a * 3 + b] = G[b * 3 + a];
    const double ratio = emia_chol_solve<3>(G, g, L, gfp);
    if (!(ratio > EMIA_ELLIPSE_FAST_MIN_RATIO)) return 0;
    for (int it = 0; it < 2; ++it) {
        double r3[3] = {0, 0, 0};
        for (int i = 0; i < P.n; ++i) {
            double px, py; emia_ellipse_pt(P, i, &px, &py);
            const double row[3] = {(px - rp[0]) * (px - rp[0]), (py - rp[1]) * (py - rp[1]), (px - rp[0]) * (py - rp[1])};
            double res = 1.0;
            for (int a = 0; a < 3; ++a) res -= row[a] * gfp[a];
            for (int a = 0; a < 3; ++a) r3[a] += row[a] * res;
        }
        double dx3[3];
        emia_chol_resolve<3>(L, r3, dx3);
        for (int a = 0; a < 3; ++a) gfp[a] += dx3[a];
    }
    return 1;
}

// General (non-"direct") fit; n >= 5 packed integer points.
EMIA_HD_NOINLINE EmiaEllipse emia_fit_ellipse_general(const uint32_t* pts, int n) {
    EmiaEllipse box; box.cx = box.cy = box.w = box.h = box.angle = 0.f; box.ok = 0;
    if (n < 5) return box;
    const double min_eps = 1e-8;
    float cxf = 0.f, cyf = 0.f;
    for (int i = 0; i < n; ++i) { cxf += (float)EMIA_PT_X(pts[i]); cyf += (float)EMIA_PT_Y(pts[i]); }
    cxf /= (float)n; cyf /= (float)n;
    double s = 0;
    for (int i = 0; i < n; ++i) {
        const float dx = (float)EMIA_PT_X(pts[i]) - cxf, dy = (float)EMIA_PT_Y(pts[i]) - cyf;
        s += fabsf(dx) + fabsf(dy);
    }
    EmiaEllipsePts P;
    P.pts = pts; P.n = n; P.cxf = cxf; P.cyf = cyf; P.perturbed = 0; P.eps = 0.f;
    P.scale = 100. / (s > FLT_EPSILON ? s : (double)FLT_EPSILON);
    const double scale = P.scale;

    double gfp[5], rp[5];
    if (!emia_ellipse_stage1_fast(P, gfp)) {
        // rotation-based path: row-streaming Givens QR; Jacobi singular values decide OpenCV's perturbation branch
        for (int attempt = 0; attempt < 2; ++attempt) {
            double R[25], qtb[5];
            for (int i = 0; i < 25; ++i) R[i] = 0.0;
            for (int i = 0; i < 5; ++i) qtb[i] = 0.0;
            for (int i = 0; i < n; ++i) {
                double px, py; emia_ellipse_pt(P, i, &px, &py);
                double row[5] = {-px * px, -py * py, -px * py, px, py};
                emia_givens_add_row<5>(R, qtb, row, 10000.0);
            }
            double smax, smin;
            emia_sv_extremes<5>(R, &smax, &smin);
            if (attempt == 0 && smax * FLT_EPSILON > smin) {
                P.eps = (float)(s / (n * 2) * 1e-3);
                P.perturbed = 1;
                continue;
            }
            emia_back_subst<5>(R, qtb, gfp);
            break;
        }
    }
    // conic centre
    {
        const double a00 = 2 * gfp[0], a01 = gfp[2], a11 = 2 * gfp[1];
        const double det = a00 * a11 - a01 * a01;
        rp[0] = (gfp[3] * a11 - a01 * gfp[4]) / det;
        rp[1] = (a00 * gfp[4] - a01 * gfp[3]) / det;
    }
    // re-fit A', B', C' about that centre
    if (!emia_ellipse_stage2_fast(P, rp, gfp)) {
        double R[9], qtb[3];
        for (int i = 0; i < 9; ++i) R[i] = 0.0;
        for (int i = 0; i < 3; ++i) qtb[i] = 0.0;
        for (int i = 0; i < n; ++i) {
            double px, py; emia_ellipse_pt(P, i, &px, &py);
            double row[3] = {(px - rp[0]) * (px - rp[0]), (py - rp[1]) * (py - rp[1]), (px - rp[0]) * (py - rp[1])};
            emia_givens_add_row<3>(R, qtb, row, 1.0);
        }
        emia_back_subst<3>(R, qtb, gfp);
    }
    double t;
    rp[4] = -0.5 * atan2(gfp[2], gfp[1] - gfp[0]);
    if (fabs(gfp[2]) > min_eps) t = gfp[2] / sin(-2.0 * rp[4]);
    else t = gfp[1] - gfp[0];
    rp[2] = fabs(gfp[0] + gfp[1] - t);
    if (rp[2] > min_eps) rp[2] = sqrt(2.0 / rp[2]);
    rp[3] = fabs(gfp[0] + gfp[1] + t);
    if (rp[3] > min_eps) rp[3] = sqrt(2.0 / rp[3]);

    box.cx = (float)(rp[0] / scale) + cxf;
    box.cy = (float)(rp[1] / scale) + cyf;
    box.w = (float)(rp[2] * 2 / scale);
    box.h = (float)(rp[3] * 2 / scale);
    if (box.w > box.h) {
        const float tmp = box.w; box.w = box.h; box.h = tmp;
        box.angle = (float)(90 + rp[4] * 180 / M_PI);
    } else {
        box.angle = (float)(rp[4] * 180 / M_PI);   // only the axes are consumed by the reference
    }
    if (box.angle < -180) box.angle += 360;
    if (box.angle > 360) box.angle -= 360;
    box.ok = 1;
    return box;
}

// cv2.fitEllipse dispatch.  OpenCV routes the exactly-determined case n == 5 to its "direct" (Halir-Flusser)
// solver; see emia_fit_ellipse_direct5 below.
EMIA_HD EmiaEllipse emia_fit_ellipse(const uint32_t* pts, int n) {
    return emia_fit_ellipse_general(pts, n);
}
