// emia_moments.cuh — image moments of a 0/1 mask up to order 3, as cv2.moments(mask.astype(np.uint8)) returns them
// (src/functions/inference.py:1101: `M = cv2.moments(mask)`; centroid = int(m10 / m00), int(m01 / m00)).  Host/device.
//
// Raw moments m_pq = sum over set pixels of x^p y^q are exact integers (< 2^53 for frames up to 8192 x 8192 and
// particle-sized masks), so they equal OpenCV's doubles bit for bit.  The SECOND-order central moments reproduce the
// arithmetic of the OpenCV 4.13 binary bit for bit (found empirically: its completeMomentState() is compiled with fused
// multiply-adds — mu20 = fma(-m10, cx, m20), mu11 = fma(-m01, cx, m11), mu02 = fma(-m01, cy, m02), cx = m10 * (1 / m00));
// the third-order central and the normalised moments follow the published formulas of completeMomentState() in plain
// double arithmetic and agree with OpenCV to rounding (they are differences of numbers ~1e6 times larger than the result:
// tests/test_hostsim_core.py states the tolerance).
#pragma once
#include "emia_common.cuh"

// order: m00 m10 m01 m20 m11 m02 m30 m21 m12 m03 | mu20 mu11 mu02 mu30 mu21 mu12 mu03 | nu20 nu11 nu02 nu30 nu21 nu12 nu03
#define EMIA_MOMENT_FIELDS 24

EMIA_HD void emia_complete_moments(const long long raw[10], double* out) {
    const double m00 = (double)raw[0], m10 = (double)raw[1], m01 = (double)raw[2], m20 = (double)raw[3], m11 = (double)raw[4],
                 m02 = (double)raw[5], m30 = (double)raw[6], m21 = (double)raw[7], m12 = (double)raw[8], m03 = (double)raw[9];
    out[0] = m00; out[1] = m10; out[2] = m01; out[3] = m20; out[4] = m11; out[5] = m02; out[6] = m30; out[7] = m21; out[8] = m12;
    out[9] = m03;
    double cx = 0, cy = 0, inv_m00 = 0.0;
    if (fabs(m00) > DBL_EPSILON) {
        inv_m00 = 1. / m00;
        cx = m10 * inv_m00;
        cy = m01 * inv_m00;
    }
    const double mu20 = emia_fma(-m10, cx, m20);
    double mu11 = emia_fma(-m01, cx, m11);
    const double mu02 = emia_fma(-m01, cy, m02);
    out[10] = mu20; out[11] = mu11; out[12] = mu02;
    out[13] = m30 - cx * (3 * mu20 + cx * m10);
    mu11 += mu11;
    out[14] = m21 - cx * (mu11 + cx * m01) - cy * mu20;
    out[15] = m12 - cy * (mu11 + cy * m10) - cx * mu02;
    out[16] = m03 - cy * (3 * mu02 + cy * m01);
    const double inv_sqrt_m00 = sqrt(fabs(inv_m00));
    const double s2 = inv_m00 * inv_m00, s3 = s2 * inv_sqrt_m00;
    out[17] = out[10] * s2; out[18] = out[11] * s2; out[19] = out[12] * s2;
    out[20] = out[13] * s3; out[21] = out[14] * s3; out[22] = out[15] * s3; out[23] = out[16] * s3;
}

// raw moments of one 32-pixel word at row y, bit 0 = pixel x0
EMIA_HD void emia_word_moments(uint32_t w, int x0, int y, long long acc[10]) {
    if (!w) return;
    long long c = 0, sx = 0, sx2 = 0, sx3 = 0;
    while (w) {
        const int b = emia_ctz(w);
        w &= w - 1;
        const long long x = x0 + b;
        c += 1; sx += x; sx2 += x * x; sx3 += x * x * x;
    }
    const long long yy = y;
    acc[0] += c; acc[1] += sx; acc[2] += c * yy; acc[3] += sx2; acc[4] += sx * yy; acc[5] += c * yy * yy;
    acc[6] += sx3; acc[7] += sx2 * yy; acc[8] += sx * yy * yy; acc[9] += c * yy * yy * yy;
}
