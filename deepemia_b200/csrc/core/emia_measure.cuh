// emia_measure.cuh — per-contour morphometry record (host/device).
//
// Replaces (reference): src/utils/measurements.py:114-233 `calculate_measurements` (contrast columns are
// computed separately from the image, :195-215) as called from the measurement loop
// src/functions/inference.py:1173-1230.
//
// Type discipline follows the reference as it executes under numpy 2.x:
//   float32 chain : box corners -> int truncation -> order_points -> midpoints -> Euclidean distances
//                   -> Length, Width, Feret_diam, Aspect_Ratio, Roundness   (um_pix rounded to float32 first)
//   float64 chain : contourArea / arcLength -> CircularED, Chords, Sphericity, Circularity (Q4: both multiplied
//                   by um_pix), ellipse axes -> major/minor axis length, eccentricity (Q3: "major" is the
//                   fitted width, i.e. the SHORTER axis).
#pragma once
#include "emia_common.cuh"
#include "emia_contour.cuh"
#include "emia_hull.cuh"
#include "emia_ellipse.cuh"

// One record per measured contour; all values stored as double (float32 results are exactly representable).
enum EmiaRecField {
    EMIA_F_MAJOR_AXIS = 0,   // "Major axis length"      (f64)
    EMIA_F_MINOR_AXIS = 1,   // "Minor axis length"      (f64)
    EMIA_F_ECCENTRICITY = 2, // "Eccentricity"           (f64)
    EMIA_F_LENGTH = 3,       // "C. Length"              (f32)
    EMIA_F_WIDTH = 4,        // "C. Width"               (f32)
    EMIA_F_CIRCULAR_ED = 5,  // "Circular eq. diameter"  (f64)
    EMIA_F_ASPECT = 6,       // "Aspect ratio"           (f32)
    EMIA_F_CIRCULARITY = 7,  // "Circularity"            (f64)
    EMIA_F_CHORDS = 8,       // "Chord length"           (f64)
    EMIA_F_FERET = 9,        // "Ferret diameter"        (f32)
    EMIA_F_ROUNDNESS = 10,   // "Roundness"              (f32)
    EMIA_F_SPHERICITY = 11,  // "Sphericity"             (f64)
    EMIA_F_AREA = 12,        // cv2.contourArea          (f64, exact)
    EMIA_F_PERIMETER = 13,   // cv2.arcLength            (f64, exact)
    EMIA_F_NVERT = 14,       // number of CHAIN_APPROX_SIMPLE vertices
    EMIA_F_RESERVED = 15,
    EMIA_F_COUNT = 16
};

// bytes of scratch needed for a contour of n vertices
EMIA_HD size_t emia_measure_scratch_bytes(int n) { return (size_t)28 * (size_t)n + 64; }
// An instance's scratch block holds one 16-byte aligned sub-block per contour, in discovery order (so that the hulls of all
// its contours can be prepared before the morphometry runs): stride of a contour of n vertices, and an upper bound of the
// block for nc contours with n_pts vertices in total.
EMIA_HD size_t emia_measure_sub_bytes(int n) { return (emia_measure_scratch_bytes(n) + 15) & ~(size_t)15; }
EMIA_HD size_t emia_measure_item_scratch_bytes(int n_pts, int nc) { return nc ? (((size_t)28 * (size_t)n_pts + (size_t)80 * (size_t)nc + 15) & ~(size_t)15) : 0; }

EMIA_HD float emia_dist2f(float ax, float ay, float bx, float by) {
    const float dx = ax - bx, dy = ay - by;
    const float dx2 = dx * dx, dy2 = dy * dy;
    return sqrtf(dx2 + dy2);
}

// imutils.perspective.order_points on 4 integer-valued float points -> tl, tr, br, bl
EMIA_HD void emia_order_points(const float* p /*8*/, float* o /*8: tl,tr,br,bl*/) {
    // stable sort of the 4 points by x (np.argsort, insertion sort for n=4)
    int idx[4] = {0, 1, 2, 3};
    for (int i = 1; i < 4; ++i) {
        const int v = idx[i];
        int j = i - 1;
        while (j >= 0 && p[2 * idx[j]] > p[2 * v]) { idx[j + 1] = idx[j]; --j; }
        idx[j + 1] = v;
    }
    int l0 = idx[0], l1 = idx[1], r0 = idx[2], r1 = idx[3];
    // left-most two sorted by y (stable)
    if (p[2 * l0 + 1] > p[2 * l1 + 1]) { const int t = l0; l0 = l1; l1 = t; }
    const int tl = l0, bl = l1;
    // right-most: np.argsort(D)[::-1] on the two distances to tl -> (br, tr); ties keep reversed order
    // (scipy cdist on the int64 corners: float64 arithmetic)
    const double ax0 = (double)p[2 * tl] - (double)p[2 * r0], ay0 = (double)p[2 * tl + 1] - (double)p[2 * r0 + 1];
    const double ax1 = (double)p[2 * tl] - (double)p[2 * r1], ay1 = (double)p[2 * tl + 1] - (double)p[2 * r1 + 1];
    const double d0 = sqrt(ax0 * ax0 + ay0 * ay0);
    const double d1 = sqrt(ax1 * ax1 + ay1 * ay1);
    // argsort ascending stable: [0,1] if d0 <= d1 else [1,0]; reversed -> first = larger (or index 1 on ties)
    int br, tr;
    if (d0 <= d1) { br = r1; tr = r0; } else { br = r0; tr = r1; }
    o[0] = p[2 * tl]; o[1] = p[2 * tl + 1];
    o[2] = p[2 * tr]; o[3] = p[2 * tr + 1];
    o[4] = p[2 * br]; o[5] = p[2 * br + 1];
    o[6] = p[2 * bl]; o[7] = p[2 * bl + 1];
}

// The three ellipse columns of a record from the fitted (width <= height) axes; zeros for contours of fewer than 5 vertices
// (src/utils/measurements.py:176-193, Q3).
EMIA_HD void emia_ellipse_columns(const EmiaEllipse& e, int n, double um_pix, double* rec) {
    double major_len = 0.0, minor_len = 0.0, ecc = 0.0;
    if (n >= 5) {
        const double major_axis = (double)e.w, minor_axis = (double)e.h;
        double a, b;
        if (major_axis > minor_axis) { a = major_axis / 2.0; b = minor_axis / 2.0; }
        else { a = minor_axis / 2.0; b = major_axis / 2.0; }
        ecc = (a != 0.0) ? sqrt(1 - ((b * b) / (a * a))) : 0.0;
        major_len = major_axis * um_pix;
        minor_len = minor_axis * um_pix;
    }
    rec[EMIA_F_MAJOR_AXIS] = major_len;
    rec[EMIA_F_MINOR_AXIS] = minor_len;
    rec[EMIA_F_ECCENTRICITY] = ecc;
}

// Where the hull kernel leaves its result inside a contour's scratch sub-block (behind the carve-up below, in the 56 spare
// bytes): int nh, then either the calipers result out[6] (nh > 2) or the first two hull points (packed, nh <= 2).
EMIA_HD size_t emia_measure_rect_off(int n) { return (size_t)28 * (size_t)n + 8; }

// Fill rec[16] for one contour.  scratch: emia_measure_scratch_bytes(n), 8-byte aligned.
// prepared: 0 = nothing, 1 = scratch starts with the n sorted hull keys, 3 = hull and rotating calipers are done (result at
// emia_measure_rect_off, written by the warp-cooperative hull kernel beforehand).
EMIA_HD_NOINLINE void emia_measure_contour(const uint32_t* pts, int n, double um_pix, void* scratch, double* rec, int prepared = 0) {
    const double area = emia_contour_area(pts, n);
    const double perimeter = emia_arc_length_closed(pts, n);

    EmiaRotRect rr;
    if (prepared == 3) {
        const int* r = (const int*)((const uint8_t*)scratch + emia_measure_rect_off(n));
        rr = emia_min_area_rect_finish(r[0], (const float*)(r + 1), (const uint32_t*)(r + 1));
    } else {
        // ---- scratch carve-up (regions are reused once dead)
        uint64_t* keys = (uint64_t*)scratch;               // 8n   (later: hull points, packed)
        int* stack = (int*)(keys + 2 * n);                 // 4(n+2), behind 8n unused bytes (layout shared with the hull kernel)
        int* hull = stack + (n + 2);                       // 4n
        int* tmp = hull + n;                               // 4n
        const int nh = emia_convex_hull(pts, n, /*clockwise=*/0, keys, stack, hull, tmp, prepared);
        uint32_t* hq = (uint32_t*)keys;
        for (int i = 0; i < nh; ++i) hq[i] = pts[hull[i]];
        rr = emia_min_area_rect_from_hull(hq, nh);
    }
    float bp[8];
    emia_box_points(rr, bp);
    for (int i = 0; i < 8; ++i) bp[i] = (float)(long long)bp[i];   // np.array(box, dtype="int"): truncation
    float o[8];
    emia_order_points(bp, o);
    const float tlx = o[0], tly = o[1], trx = o[2], try_ = o[3], brx = o[4], bry = o[5], blx = o[6], bly = o[7];
    const float tltrX = (tlx + trx) * 0.5f, tltrY = (tly + try_) * 0.5f;
    const float blbrX = (blx + brx) * 0.5f, blbrY = (bly + bry) * 0.5f;
    const float tlblX = (tlx + blx) * 0.5f, tlblY = (tly + bly) * 0.5f;
    const float trbrX = (trx + brx) * 0.5f, trbrY = (try_ + bry) * 0.5f;
    const float dA = emia_dist2f(tltrX, tltrY, blbrX, blbrY);
    const float dB = emia_dist2f(tlblX, tlblY, trbrX, trbrY);
    const float umf = (float)um_pix;
    const float dmax = dA > dB ? dA : dB;   // python max(dimA, dimB): returns dimA unless dimB > dimA
    const float dmin = dB < dA ? dB : dA;   // python min(dimA, dimB)
    float aspect = 0.f;
    if (dA != 0.f && dB != 0.f) aspect = dmax / dmin;
    rec[EMIA_F_LENGTH] = (double)(dmin * umf);
    rec[EMIA_F_WIDTH] = (double)(dmax * umf);
    rec[EMIA_F_FERET] = (double)(dmax * umf);
    rec[EMIA_F_ASPECT] = (double)aspect;
    rec[EMIA_F_ROUNDNESS] = (aspect != 0.f) ? (double)(1.0f / aspect) : 0.0;

    rec[EMIA_F_CIRCULAR_ED] = sqrt(4 * area / M_PI) * um_pix;
    rec[EMIA_F_CHORDS] = perimeter * um_pix;
    rec[EMIA_F_SPHERICITY] = (perimeter != 0.0) ? (2 * sqrt(M_PI * area)) / perimeter * um_pix : 0.0;
    rec[EMIA_F_CIRCULARITY] = (perimeter != 0.0) ? (4 * M_PI) * (area / (perimeter * perimeter)) * um_pix : 0.0;

    EmiaEllipse e; e.w = e.h = 0.f;
    if (n >= 5) e = emia_fit_ellipse(pts, n);
    emia_ellipse_columns(e, n, um_pix, rec);
    rec[EMIA_F_AREA] = area;
    rec[EMIA_F_PERIMETER] = perimeter;
    rec[EMIA_F_NVERT] = (double)n;
    rec[EMIA_F_RESERVED] = 0.0;
}
