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

// Minimum-norm least-squares solution from the QR factor: R x = qtb with R = U S V' by one-sided (Hestenes) Jacobi on a copy of R;
// x = sum over singular values s_j > 2 * DBL_EPSILON * sum(s) of V_j (U_j' qtb) / s_j — what cv::SVD::backSubst / cv::solve(
// DECOMP_SVD) return, including for rank-deficient systems (e.g. stage 2 when |dx| == |dy| for every vertex: the x^2 and y^2
// columns coincide and OpenCV answers with the equal split, a circle).  smax / smin: extreme singular values.
template <int K>
EMIA_HD_COLD void emia_svd_solve(const double* R, const double* qtb, double* x, double* smax, double* smin) {
    double M[K * K], V[K * K];
    for (int i = 0; i < K * K; ++i) { M[i] = R[i]; V[i] = 0.0; }
    for (int i = 0; i < K; ++i) V[i * K + i] = 1.0;
    for (int sweep = 0; sweep < 40; ++sweep) {
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
                const double c = 1.0 / sqrt(1.0 + t * t), sn = c * t;
                for (int i = 0; i < K; ++i) {
                    const double mp = M[i * K + p], mq = M[i * K + q];
                    M[i * K + p] = c * mp - sn * mq;
                    M[i * K + q] = sn * mp + c * mq;
                    const double vp = V[i * K + p], vq = V[i * K + q];
                    V[i * K + p] = c * vp - sn * vq;
                    V[i * K + q] = sn * vp + c * vq;
                }
            }
        if (off < 1e-16) break;
    }
    double sv[K], sum = 0.0, mx = 0.0, mn = 1e300;
    for (int j = 0; j < K; ++j) {
        double nrm = 0.0;
        for (int i = 0; i < K; ++i) nrm += M[i * K + j] * M[i * K + j];
        sv[j] = sqrt(nrm);
        sum += sv[j];
        if (sv[j] > mx) mx = sv[j];
        if (sv[j] < mn) mn = sv[j];
    }
    const double thr = sum * (2.0 * DBL_EPSILON);
    for (int i = 0; i < K; ++i) x[i] = 0.0;
    for (int j = 0; j < K; ++j) {
        if (!(sv[j] > thr)) continue;
        double proj = 0.0;                       // (U_j s_j)' qtb
        for (int i = 0; i < K; ++i) proj += M[i * K + j] * qtb[i];
        const double coef = proj / (sv[j] * sv[j]);
        for (int i = 0; i < K; ++i) x[i] += coef * V[i * K + j];
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
// q = P.pts[i], loaded by the caller one iteration ahead (the loops below are software-pipelined by hand: the load of vertex
// i + 1 is in flight while the ~20 dependent fp64 operations of vertex i execute)
EMIA_HD void emia_ellipse_pt(const EmiaEllipsePts& P, int i, uint32_t q, double* px, double* py) {
    float fx = (float)EMIA_PT_X(q), fy = (float)EMIA_PT_Y(q);
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
    uint32_t qn = P.pts[0];
    for (int i = 0; i < P.n; ++i) {
        const uint32_t q = qn;
        if (i + 1 < P.n) qn = P.pts[i + 1];
        double px, py; emia_ellipse_pt(P, i, q, &px, &py);
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
        uint32_t qn = P.pts[0];
        for (int i = 0; i < P.n; ++i) {
            const uint32_t q = qn;
            if (i + 1 < P.n) qn = P.pts[i + 1];
            double px, py; emia_ellipse_pt(P, i, q, &px, &py);
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
    uint32_t qn = P.pts[0];
    for (int i = 0; i < P.n; ++i) {
        const uint32_t q = qn;
        if (i + 1 < P.n) qn = P.pts[i + 1];
        double px, py; emia_ellipse_pt(P, i, q, &px, &py);
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
        uint32_t qn = P.pts[0];
        for (int i = 0; i < P.n; ++i) {
            const uint32_t q = qn;
            if (i + 1 < P.n) qn = P.pts[i + 1];
            double px, py; emia_ellipse_pt(P, i, q, &px, &py);
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

// minimum-norm solution of the symmetric system [a b; b c] x = (r0, r1) through its eigen-decomposition (|eigenvalues| are the
// singular values); components whose singular value is <= 2 * DBL_EPSILON * (s1 + s2) are dropped, as cv::SVD::backSubst does.
EMIA_HD void emia_solve_sym2_minnorm(double a, double b, double c, double r0, double r1, double* x0, double* x1) {
    const double half_tr = 0.5 * (a + c), half_diff = 0.5 * (a - c);
    const double rad = sqrt(half_diff * half_diff + b * b);
    const double l1 = half_tr + rad, l2 = half_tr - rad;
    // eigenvector of l1: (b, l1 - a) or (l1 - c, b), whichever is better conditioned; l2's is orthogonal
    double vx, vy;
    if (fabs(l1 - a) > fabs(l1 - c)) { vx = b; vy = l1 - a; } else { vx = l1 - c; vy = b; }
    double nrm = sqrt(vx * vx + vy * vy);
    if (nrm == 0.0) { vx = 1.0; vy = 0.0; nrm = 1.0; }     // a == c, b == 0: any basis
    vx /= nrm; vy /= nrm;
    const double thr = (fabs(l1) + fabs(l2)) * (2.0 * DBL_EPSILON);
    double sx = 0.0, sy = 0.0;
    if (fabs(l1) > thr) { const double t = (vx * r0 + vy * r1) / l1; sx += t * vx; sy += t * vy; }
    if (fabs(l2) > thr) { const double t = (-vy * r0 + vx * r1) / l2; sx += t * -vy; sy += t * vx; }
    *x0 = sx; *x1 = sy;
}

// Axes / angle of the re-fitted conic A' x^2 + B' y^2 + C' xy = 1 about the centre rp[0..1] (scaled coordinates), rounded to
// float32 as OpenCV does.  rp[2..4] are work space.  Shared by the serial fit and the cooperative kernel.
EMIA_HD EmiaEllipse emia_ellipse_from_conic(const double* gfp, double* rp, double scale, float cxf, float cyf) {
    EmiaEllipse box;
    const double min_eps = 1e-8;
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

// General (non-"direct") fit; n >= 5 packed integer points.
EMIA_HD_NOINLINE EmiaEllipse emia_fit_ellipse_general(const uint32_t* pts, int n) {
    EmiaEllipse box; box.cx = box.cy = box.w = box.h = box.angle = 0.f; box.ok = 0;
    if (n < 5) return box;
    float cxf = 0.f, cyf = 0.f;
#pragma unroll 4
    for (int i = 0; i < n; ++i) { cxf += (float)EMIA_PT_X(pts[i]); cyf += (float)EMIA_PT_Y(pts[i]); }
    cxf /= (float)n; cyf /= (float)n;
    double s = 0;
#pragma unroll 4
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
                double px, py; emia_ellipse_pt(P, i, P.pts[i], &px, &py);
                double row[5] = {-px * px, -py * py, -px * py, px, py};
                emia_givens_add_row<5>(R, qtb, row, 10000.0);
            }
            double smax, smin;
            emia_svd_solve<5>(R, qtb, gfp, &smax, &smin);
            if (attempt == 0 && smax * FLT_EPSILON > smin) {
                // OpenCV re-fits perturbed points here (4.13 draws the offsets from an RNG, so its own result is not
                // reproducible on such inputs); a deterministic offset pattern stands in
                P.eps = (float)(s / (n * 2) * 1e-3);
                P.perturbed = 1;
                continue;
            }
            break;
        }
    }
    // conic centre: OpenCV solves [2A C; C 2B] r = [D; E] with cv::solve(DECOMP_SVD), i.e. the MINIMUM-NORM solution with
    // singular values <= 2 eps (s1 + s2) dropped.  Degenerate conics (two parallel lines: a mask two pixels wide) make this
    // matrix singular, so the plain 2x2 inverse is not a substitute.
    emia_solve_sym2_minnorm(2 * gfp[0], gfp[2], 2 * gfp[1], gfp[3], gfp[4], &rp[0], &rp[1]);
    // re-fit A', B', C' about that centre
    if (!emia_ellipse_stage2_fast(P, rp, gfp)) {
        double R[9], qtb[3];
        for (int i = 0; i < 9; ++i) R[i] = 0.0;
        for (int i = 0; i < 3; ++i) qtb[i] = 0.0;
        for (int i = 0; i < n; ++i) {
            double px, py; emia_ellipse_pt(P, i, P.pts[i], &px, &py);
            double row[3] = {(px - rp[0]) * (px - rp[0]), (py - rp[1]) * (py - rp[1]), (px - rp[0]) * (py - rp[1])};
            emia_givens_add_row<3>(R, qtb, row, 1.0);
        }
        double smax, smin;
        emia_svd_solve<3>(R, qtb, gfp, &smax, &smin);
    }
    return emia_ellipse_from_conic(gfp, rp, scale, cxf, cyf);
}

// ---- n == 5: OpenCV's fitEllipse routes the exactly-determined case to fitEllipseDirect (Halir & Flusser, "Numerically stable
// direct least squares fitting of ellipses"): with D1 = [x^2 xy y^2], D2 = [x y 1] (points centred on their mean and scaled by
// 100 / sum(|dx| + |dy|)), S1 = D1'D1, S2 = D1'D2, S3 = D2'D2, T = -S3^-1 S2', M = C1^-1 (S1 + S2 T); the conic (A,B,C) is the
// eigenvector of M with 4AC - B^2 > 0 and (D,E,F) = T (A,B,C).  The solution is unique, so any stable evaluation agrees with
// OpenCV's closed-form one to rounding.  (OpenCV additionally re-fits RANDOMLY perturbed points when |det M| <= 1e-10; for five
// points det M is pure rounding noise, so that branch is taken on ~3 % of inputs and makes cv2.fitEllipse itself
// non-reproducible there — call it twice, get two answers.  This implementation always returns the unperturbed solution.)
EMIA_HD void emia_mat3_inv(const double* m, double* inv, double* det_out) {
    const double c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
    const double det = m[0] * c00 + m[1] * c01 + m[2] * c02;
    const double id = 1.0 / det;
    inv[0] = c00 * id; inv[1] = (m[2] * m[7] - m[1] * m[8]) * id; inv[2] = (m[1] * m[5] - m[2] * m[4]) * id;
    inv[3] = c01 * id; inv[4] = (m[0] * m[8] - m[2] * m[6]) * id; inv[5] = (m[2] * m[3] - m[0] * m[5]) * id;
    inv[6] = c02 * id; inv[7] = (m[1] * m[6] - m[0] * m[7]) * id; inv[8] = (m[0] * m[4] - m[1] * m[3]) * id;
    *det_out = det;
}
// real eigenvalues of a 3x3 matrix with real spectrum (trigonometric solution of the characteristic cubic)
EMIA_HD void emia_mat3_eigvals(const double* m, double* ev) {
    const double tr = m[0] + m[4] + m[8];
    const double c1 = m[0] * m[4] - m[1] * m[3] + m[0] * m[8] - m[2] * m[6] + m[4] * m[8] - m[5] * m[7];
    const double det = m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
    // lambda = tr/3 + t,  t^3 + p t + q = 0
    const double sh = tr / 3.0;
    const double pp = c1 - tr * tr / 3.0;
    const double qq = -2.0 * tr * tr * tr / 27.0 + tr * c1 / 3.0 - det;
    if (pp >= 0.0) { ev[0] = ev[1] = ev[2] = sh; return; }
    const double r = 2.0 * sqrt(-pp / 3.0);
    double arg = 3.0 * qq / (pp * r);
    arg = arg < -1.0 ? -1.0 : (arg > 1.0 ? 1.0 : arg);
    const double phi = acos(arg) / 3.0;
    for (int k = 0; k < 3; ++k) ev[k] = sh + r * cos(phi - 2.0 * M_PI * k / 3.0);
}
// null vector of (M - lambda I): the largest cross product of two of its rows
EMIA_HD void emia_mat3_eigvec(const double* m, double lambda, double* v) {
    const double a[9] = {m[0] - lambda, m[1], m[2], m[3], m[4] - lambda, m[5], m[6], m[7], m[8] - lambda};
    double best = -1.0;
    for (int i = 0; i < 3; ++i) {
        const double* r0 = a + 3 * i;
        const double* r1 = a + 3 * ((i + 1) % 3);
        const double cx = r0[1] * r1[2] - r0[2] * r1[1], cy = r0[2] * r1[0] - r0[0] * r1[2], cz = r0[0] * r1[1] - r0[1] * r1[0];
        const double nn = cx * cx + cy * cy + cz * cz;
        if (nn > best) { best = nn; v[0] = cx; v[1] = cy; v[2] = cz; }
    }
    const double nrm = sqrt(best);
    if (nrm > 0.0) { v[0] /= nrm; v[1] /= nrm; v[2] /= nrm; }
}
EMIA_HD_COLD EmiaEllipse emia_fit_ellipse_direct(const uint32_t* pts, int n) {
    EmiaEllipse box; box.cx = box.cy = box.w = box.h = box.angle = 0.f; box.ok = 0;
    double cx = 0.0, cy = 0.0;
    for (int i = 0; i < n; ++i) { cx += (double)EMIA_PT_X(pts[i]); cy += (double)EMIA_PT_Y(pts[i]); }
    cx /= (double)n; cy /= (double)n;
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += fabs((double)EMIA_PT_X(pts[i]) - cx) + fabs((double)EMIA_PT_Y(pts[i]) - cy);
    const double scale = 100. / (s > FLT_EPSILON ? s : (double)FLT_EPSILON);
    double S1[9] = {0}, S2[9] = {0}, S3[9] = {0};
    for (int i = 0; i < n; ++i) {
        const double px = ((double)EMIA_PT_X(pts[i]) - cx) * scale, py = ((double)EMIA_PT_Y(pts[i]) - cy) * scale;
        const double d1[3] = {px * px, px * py, py * py}, d2[3] = {px, py, 1.0};
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) { S1[a * 3 + b] += d1[a] * d1[b]; S2[a * 3 + b] += d1[a] * d2[b]; S3[a * 3 + b] += d2[a] * d2[b]; }
    }
    const double inv_n = 1.0 / (double)n;
    for (int k = 0; k < 9; ++k) { S1[k] *= inv_n; S2[k] *= inv_n; S3[k] *= inv_n; }
    double S3i[9], det3;
    emia_mat3_inv(S3, S3i, &det3);
    if (!(fabs(det3) > 0.0)) return box;
    double T[9];      // T = -S3^-1 S2'
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) {
            double v = 0.0;
            for (int k = 0; k < 3; ++k) v += S3i[a * 3 + k] * S2[b * 3 + k];
            T[a * 3 + b] = -v;
        }
    double M0[9];     // S1 + S2 T
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) {
            double v = S1[a * 3 + b];
            for (int k = 0; k < 3; ++k) v += S2[a * 3 + k] * T[k * 3 + b];
            M0[a * 3 + b] = v;
        }
    double M[9];      // C1^-1 M0: rows (M0[2] / 2, -M0[1], M0[0] / 2)
    for (int b = 0; b < 3; ++b) { M[b] = 0.5 * M0[6 + b]; M[3 + b] = -M0[3 + b]; M[6 + b] = 0.5 * M0[b]; }
    double ev[3];
    emia_mat3_eigvals(M, ev);
    double a1[3] = {0, 0, 0}, best_cond = 0.0;
    for (int k = 0; k < 3; ++k) {
        double v[3];
        emia_mat3_eigvec(M, ev[k], v);
        const double cond = 4.0 * v[0] * v[2] - v[1] * v[1];
        if (cond > best_cond) { best_cond = cond; a1[0] = v[0]; a1[1] = v[1]; a1[2] = v[2]; }
    }
    if (!(best_cond > 0.0)) return box;
    const double A = a1[0], B = a1[1], C = a1[2];
    const double D = T[0] * A + T[1] * B + T[2] * C, E = T[3] * A + T[4] * B + T[5] * C, F = T[6] * A + T[7] * B + T[8] * C;
    const double den = B * B - 4.0 * A * C;
    const double x0 = (2.0 * C * D - B * E) / den, y0 = (2.0 * A * E - B * D) / den;
    const double num = 2.0 * (A * E * E + C * D * D - B * D * E + den * F);
    const double l1 = sqrt((A - C) * (A - C) + B * B);
    const double ra = -sqrt(num * ((A + C) + l1)) / den, rb = -sqrt(num * ((A + C) - l1)) / den;
    if (!(ra == ra) || !(rb == rb)) return box;        // NaN: no real ellipse
    box.cx = (float)(x0 / scale + cx);
    box.cy = (float)(y0 / scale + cy);
    box.w = (float)(2.0 * ra / scale);
    box.h = (float)(2.0 * rb / scale);
    if (box.w > box.h) { const float t = box.w; box.w = box.h; box.h = t; }
    box.angle = 0.f;                                    // only the axes are consumed by the reference
    box.ok = 1;
    return box;
}

// cv2.fitEllipse dispatch: n == 5 -> direct fit (falls back to the general fit when it finds no ellipse), else the general fit.
EMIA_HD EmiaEllipse emia_fit_ellipse(const uint32_t* pts, int n) {
    if (n == 5) {
        const EmiaEllipse d = emia_fit_ellipse_direct(pts, n);
        if (d.ok) return d;
    }
    return emia_fit_ellipse_general(pts, n);
}
