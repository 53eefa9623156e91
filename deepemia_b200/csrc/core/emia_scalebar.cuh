// emia_scalebar.cuh — row f3: the line-detection part of the reference's scale-bar detector
// (/root/reference/src/utils/scalebar_ocr.py:72-373): cv2.cvtColor(BGR2GRAY) -> cv2.Canny(50, 150, apertureSize=3) ->
// cv2.HoughLinesP(rho 1, theta pi/180, threshold 50, minLineLength 20, maxLineGap 10) -> per line cv2.line(mask, thickness 2) +
// cv2.mean(gray, mask).  The third-party arithmetic (OpenCV 4.x imgproc: canny.cpp, hough.cpp HoughLinesProbabilistic,
// drawing.cpp ThickLine / FillConvexPoly / Line2 / Circle, core RNG) is restated here so that the results are bit-identical to
// OpenCV's; the OCR (EasyOCR, a CRNN) stays outside.  EMIA_HD: the same source is compiled by g++ into
// tests/hostsim and checked against cv2 on the CPU.
#pragma once
#include "emia_common.cuh"

// ------------------------------------------------------------------------------------------------------------------
// cv::RNG (multiply-with-carry), cvRound
// ------------------------------------------------------------------------------------------------------------------
EMIA_HD uint32_t emia_cv_rng_next(uint64_t& state) {
    state = (uint64_t)(uint32_t)state * 4164903690ULL + (uint32_t)(state >> 32);
    return (uint32_t)state;
}
// RNG::uniform(0, count)
EMIA_HD int emia_cv_rng_uniform0(uint64_t& state, int count) { return count == 0 ? 0 : (int)(emia_cv_rng_next(state) % (uint32_t)count); }

EMIA_HD int emia_cv_round_f(float v) {
#if defined(__CUDA_ARCH__)
    return __float2int_rn(v);
#else
    return (int)lrintf(v);
#endif
}
EMIA_HD int emia_cv_round_d(double v) {
#if defined(__CUDA_ARCH__)
    return __double2int_rn(v);
#else
    return (int)lrint(v);
#endif
}

// ------------------------------------------------------------------------------------------------------------------
// cv2.cvtColor(BGR2GRAY), 8 bit (OpenCV 4.13: 15-bit coefficients)
// ------------------------------------------------------------------------------------------------------------------
EMIA_HD uint8_t emia_bgr2gray(int b, int g, int r) { return (uint8_t)((b * 3735 + g * 19235 + r * 9798 + 16384) >> 15); }

// ------------------------------------------------------------------------------------------------------------------
// cv2.Canny(gray, low, high, apertureSize=3, L2gradient=False): Sobel 3x3 with BORDER_REPLICATE, |dx| + |dy|,
// non-maximum suppression on the fixed-point tan(22.5) sectors, double threshold.  g: H x W bytes, row pitch `pitch`.
// ------------------------------------------------------------------------------------------------------------------
EMIA_HD int emia_px_rep(const uint8_t* g, int H, int W, int pitch, int x, int y) {
    x = x < 0 ? 0 : (x >= W ? W - 1 : x);
    y = y < 0 ? 0 : (y >= H ? H - 1 : y);
    return g[(size_t)y * pitch + x];
}
EMIA_HD void emia_sobel3(const uint8_t* g, int H, int W, int pitch, int x, int y, int& dx, int& dy) {
    int a = emia_px_rep(g, H, W, pitch, x - 1, y - 1), b = emia_px_rep(g, H, W, pitch, x, y - 1), c = emia_px_rep(g, H, W, pitch, x + 1, y - 1);
    int d = emia_px_rep(g, H, W, pitch, x - 1, y), f = emia_px_rep(g, H, W, pitch, x + 1, y);
    int p = emia_px_rep(g, H, W, pitch, x - 1, y + 1), q = emia_px_rep(g, H, W, pitch, x, y + 1), r = emia_px_rep(g, H, W, pitch, x + 1, y + 1);
    dx = (c + 2 * f + r) - (a + 2 * d + p);
    dy = (p + 2 * q + r) - (a + 2 * b + c);
}
// gradient magnitude; zero outside the image (OpenCV pads the magnitude rows / columns with zeros)
EMIA_HD int emia_canny_mag(const uint8_t* g, int H, int W, int pitch, int x, int y) {
    if ((unsigned)x >= (unsigned)W || (unsigned)y >= (unsigned)H) return 0;
    int dx, dy;
    emia_sobel3(g, H, W, pitch, x, y, dx, dy);
    return (dx < 0 ? -dx : dx) + (dy < 0 ? -dy : dy);
}
// 2: strong edge, 0: weak candidate (kept only when connected to a strong one), 1: no edge
EMIA_HD int emia_canny_classify(const uint8_t* g, int H, int W, int pitch, int x, int y, int low, int high) {
    int xs, ys;
    emia_sobel3(g, H, W, pitch, x, y, xs, ys);
    int ax = xs < 0 ? -xs : xs, ay = ys < 0 ? -ys : ys;
    int m = ax + ay;
    if (m <= low) return 1;
    const int TG22 = 13573;                       // (int)(0.41421356237309504 * (1 << 15) + 0.5)
    int yy = ay << 15;
    int tg22x = ax * TG22;
    bool keep;
    if (yy < tg22x) {
        keep = m > emia_canny_mag(g, H, W, pitch, x - 1, y) && m >= emia_canny_mag(g, H, W, pitch, x + 1, y);
    } else {
        int tg67x = tg22x + (ax << 16);
        if (yy > tg67x) {
            keep = m > emia_canny_mag(g, H, W, pitch, x, y - 1) && m >= emia_canny_mag(g, H, W, pitch, x, y + 1);
        } else {
            int s = ((xs ^ ys) < 0) ? -1 : 1;
            keep = m > emia_canny_mag(g, H, W, pitch, x - s, y - 1) && m > emia_canny_mag(g, H, W, pitch, x + s, y + 1);
        }
    }
    if (!keep) return 1;
    return m > high ? 2 : 0;
}

// ------------------------------------------------------------------------------------------------------------------
// cv2.HoughLinesP pieces (hough.cpp, HoughLinesProbabilistic).  trig[2n] = (float)(cos(n*theta) / rho), trig[2n+1] = sin.
// ------------------------------------------------------------------------------------------------------------------
EMIA_HD int emia_hough_rho_bin(int j, int i, float c, float s, int numrho) {
    float v = (float)j * c + (float)i * s;        // never contracted (-fmad=false / -ffp-contract=off)
    int r = emia_cv_round_f(v) + (numrho - 1) / 2;
    return r < 0 ? 0 : (r >= numrho ? numrho - 1 : r);        // no effect with OpenCV's numrho; keeps a caller's bad table in bounds
}

struct EmiaHoughWalk {
    int x0, y0, dx0, dy0, xflag;
};
// direction of the most voted line through (j, i): fixed-point (16 bit) stepping along the dominant axis
EMIA_HD EmiaHoughWalk emia_hough_walk_setup(int j, int i, float cos_n, float sin_n) {
    EmiaHoughWalk w;
    float a = -sin_n, b = cos_n;
    w.x0 = j; w.y0 = i;
    if (fabsf(a) > fabsf(b)) {
        w.xflag = 1;
        w.dx0 = a > 0 ? 1 : -1;
        w.dy0 = emia_cv_round_f(b * 65536.0f / fabsf(a));            // float arithmetic, as OpenCV's expression is
        w.y0 = (w.y0 << 16) + (1 << 15);
    } else {
        w.xflag = 0;
        w.dy0 = b > 0 ? 1 : -1;
        w.dx0 = emia_cv_round_f(a * 65536.0f / fabsf(b));
        w.x0 = (w.x0 << 16) + (1 << 15);
    }
    return w;
}
// pixel visited at step t (t >= 0) in direction k (0: forward, 1: backward)
EMIA_HD void emia_hough_walk_at(const EmiaHoughWalk& w, int k, int t, int& j1, int& i1) {
    int dx = k ? -w.dx0 : w.dx0, dy = k ? -w.dy0 : w.dy0;
    int x = w.x0 + t * dx, y = w.y0 + t * dy;
    if (w.xflag) { j1 = x; i1 = y >> 16; }
    else { j1 = x >> 16; i1 = y; }
}

// ------------------------------------------------------------------------------------------------------------------
// cv2.line(mask, (x1,y1), (x2,y2), 255, thickness=2) (drawing.cpp: ThickLine -> FillConvexPoly + Line2 outline + two
// radius-1 filled circles).  `put(x, y)` is called for every covered pixel inside [0,W) x [0,H) (possibly more than once).
// ------------------------------------------------------------------------------------------------------------------
#define EMIA_XY_SHIFT 16
#define EMIA_XY_ONE (1 << EMIA_XY_SHIFT)

struct EmiaPt64 {
    int64_t x, y;
};

EMIA_HD bool emia_cv_clip_line(int64_t width, int64_t height, EmiaPt64& p1, EmiaPt64& p2) {
    int64_t right = width - 1, bottom = height - 1;
    if (width <= 0 || height <= 0) return false;
    int64_t &x1 = p1.x, &y1 = p1.y, &x2 = p2.x, &y2 = p2.y;
    int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
    int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
        int64_t a;
        if (c1 & 12) {
            a = c1 < 8 ? 0 : bottom;
            x1 += (int64_t)((double)(a - y1) * (double)(x2 - x1) / (double)(y2 - y1));
            y1 = a;
            c1 = (x1 < 0) + (x1 > right) * 2;
        }
        if (c2 & 12) {
            a = c2 < 8 ? 0 : bottom;
            x2 += (int64_t)((double)(a - y2) * (double)(x2 - x1) / (double)(y2 - y1));
            y2 = a;
            c2 = (x2 < 0) + (x2 > right) * 2;
        }
        if ((c1 & c2) == 0 && (c1 | c2) != 0) {
            if (c1) {
                a = c1 == 1 ? 0 : right;
                y1 += (int64_t)((double)(a - x1) * (double)(y2 - y1) / (double)(x2 - x1));
                x1 = a;
                c1 = 0;
            }
            if (c2) {
                a = c2 == 1 ? 0 : right;
                y2 += (int64_t)((double)(a - x2) * (double)(y2 - y1) / (double)(x2 - x1));
                x2 = a;
                c2 = 0;
            }
        }
    }
    return (c1 | c2) == 0;
}

// Line2: the 8-connected line between two fixed-point (16 fractional bits) end points
template <class Put>
EMIA_HD void emia_cv_line2(int W, int H, EmiaPt64 pt1, EmiaPt64 pt2, Put put) {
    if (!emia_cv_clip_line((int64_t)W << EMIA_XY_SHIFT, (int64_t)H << EMIA_XY_SHIFT, pt1, pt2)) return;
    int64_t dx = pt2.x - pt1.x, dy = pt2.y - pt1.y;
    int64_t j = dx < 0 ? -1 : 0;
    int64_t ax = (dx ^ j) - j;
    int64_t i = dy < 0 ? -1 : 0;
    int64_t ay = (dy ^ i) - i;
    int64_t x_step, y_step;
    int ecount;
    if (ax > ay) {
        dy = (dy ^ j) - j;
        pt1.x ^= pt2.x & j; pt2.x ^= pt1.x & j; pt1.x ^= pt2.x & j;
        pt1.y ^= pt2.y & j; pt2.y ^= pt1.y & j; pt1.y ^= pt2.y & j;
        x_step = EMIA_XY_ONE;
        y_step = (dy * (int64_t)EMIA_XY_ONE) / (ax | 1);
        ecount = (int)((pt2.x - pt1.x) >> EMIA_XY_SHIFT);
    } else {
        dx = (dx ^ i) - i;
        pt1.x ^= pt2.x & i; pt2.x ^= pt1.x & i; pt1.x ^= pt2.x & i;
        pt1.y ^= pt2.y & i; pt2.y ^= pt1.y & i; pt1.y ^= pt2.y & i;
        x_step = (dx * (int64_t)EMIA_XY_ONE) / (ay | 1);
        y_step = EMIA_XY_ONE;
        ecount = (int)((pt2.y - pt1.y) >> EMIA_XY_SHIFT);
    }
    pt1.x += (EMIA_XY_ONE >> 1);
    pt1.y += (EMIA_XY_ONE >> 1);
    {
        int x = (int)((pt2.x + (EMIA_XY_ONE >> 1)) >> EMIA_XY_SHIFT), y = (int)((pt2.y + (EMIA_XY_ONE >> 1)) >> EMIA_XY_SHIFT);
        if (0 <= x && x < W && 0 <= y && y < H) put(x, y);
    }
    if (ax > ay) {
        pt1.x >>= EMIA_XY_SHIFT;
        while (ecount >= 0) {
            int x = (int)pt1.x, y = (int)(pt1.y >> EMIA_XY_SHIFT);
            if (0 <= x && x < W && 0 <= y && y < H) put(x, y);
            pt1.x++;
            pt1.y += y_step;
            ecount--;
        }
    } else {
        pt1.y >>= EMIA_XY_SHIFT;
        while (ecount >= 0) {
            int x = (int)(pt1.x >> EMIA_XY_SHIFT), y = (int)pt1.y;
            if (0 <= x && x < W && 0 <= y && y < H) put(x, y);
            pt1.x += x_step;
            pt1.y++;
            ecount--;
        }
    }
    (void)x_step;
}

// FillConvexPoly(v[4], line_type 8, shift XY_SHIFT): outline with Line2, then scan-line fill between the two edge chains
template <class Put>
EMIA_HD void emia_cv_fill_convex4(int W, int H, const EmiaPt64* v, Put put) {
    const int npts = 4, shift = EMIA_XY_SHIFT;
    struct Edge { int idx, di; int64_t x, dx; int ye; } edge[2];
    const int delta = 1 << shift >> 1;
    int imin = 0, edges = npts;
    int64_t xmin = v[0].x, xmax = v[0].x, ymin = v[0].y, ymax = v[0].y;
    const int delta1 = EMIA_XY_ONE >> 1, delta2 = EMIA_XY_ONE >> 1;
    EmiaPt64 p0 = v[npts - 1];
    for (int i = 0; i < npts; i++) {
        EmiaPt64 p = v[i];
        if (p.y < ymin) { ymin = p.y; imin = i; }
        ymax = ymax > p.y ? ymax : p.y;
        xmax = xmax > p.x ? xmax : p.x;
        xmin = xmin < p.x ? xmin : p.x;
        emia_cv_line2(W, H, p0, p, put);
        p0 = p;
    }
    xmin = (xmin + delta) >> shift; xmax = (xmax + delta) >> shift;
    ymin = (ymin + delta) >> shift; ymax = (ymax + delta) >> shift;
    if ((int)xmax < 0 || (int)ymax < 0 || (int)xmin >= W || (int)ymin >= H) return;
    ymax = ymax < H - 1 ? ymax : H - 1;
    edge[0].idx = edge[1].idx = imin;
    int y = (int)ymin;
    edge[0].ye = edge[1].ye = y;
    edge[0].di = 1; edge[1].di = npts - 1;
    edge[0].x = edge[1].x = -(int64_t)EMIA_XY_ONE;
    edge[0].dx = edge[1].dx = 0;
    do {
        for (int i = 0; i < 2; i++) {
            if (y >= edge[i].ye) {
                int idx0 = edge[i].idx, di = edge[i].di;
                int idx = idx0 + di;
                if (idx >= npts) idx -= npts;
                int ty = 0;
                for (; edges-- > 0;) {
                    ty = (int)((v[idx].y + delta) >> shift);
                    if (ty > y) {
                        int64_t xs = v[idx0].x, xe = v[idx].x;
                        edge[i].ye = ty;
                        edge[i].dx = ((xe - xs) * 2 + (ty - y)) / (2 * (ty - y));
                        edge[i].x = xs;
                        edge[i].idx = idx;
                        break;
                    }
                    idx0 = idx;
                    idx += di;
                    if (idx >= npts) idx -= npts;
                }
            }
        }
        if (edges < 0) break;
        if (y >= 0) {
            int left = 0, right = 1;
            if (edge[0].x > edge[1].x) { left = 1; right = 0; }
            int xx1 = (int)((edge[left].x + delta1) >> EMIA_XY_SHIFT);
            int xx2 = (int)((edge[right].x + delta2) >> EMIA_XY_SHIFT);
            if (xx2 >= 0 && xx1 < W) {
                if (xx1 < 0) xx1 = 0;
                if (xx2 >= W) xx2 = W - 1;
                for (int x = xx1; x <= xx2; ++x) put(x, y);
            }
        }
        edge[0].x += edge[0].dx;
        edge[1].x += edge[1].dx;
    } while (++y <= (int)ymax);
}

template <class Put>
EMIA_HD void emia_cv_plus1(int W, int H, int cx, int cy, Put put) {          // Circle(center, radius 1, filled)
    for (int k = 0; k < 5; ++k) {
        int x = cx + (k == 1) - (k == 2), y = cy + (k == 3) - (k == 4);
        if (0 <= x && x < W && 0 <= y && y < H) put(x, y);
    }
}

template <class Put>
EMIA_HD void emia_cv_thick_line2(int W, int H, int x1, int y1, int x2, int y2, Put put) {
    EmiaPt64 p0{(int64_t)x1 << EMIA_XY_SHIFT, (int64_t)y1 << EMIA_XY_SHIFT}, p1{(int64_t)x2 << EMIA_XY_SHIFT, (int64_t)y2 << EMIA_XY_SHIFT};
    const double inv = 1.0 / EMIA_XY_ONE;
    double dx = (double)(p0.x - p1.x) * inv, dy = (double)(p1.y - p0.y) * inv;
    double r = dx * dx + dy * dy;
    const int thickness = 2 << (EMIA_XY_SHIFT - 1);
    if (fabs(r) > DBL_EPSILON) {
        r = (double)thickness / sqrt(r);
        int64_t dpx = emia_cv_round_d(dy * r), dpy = emia_cv_round_d(dx * r);
        EmiaPt64 pt[4] = {{p0.x + dpx, p0.y + dpy}, {p0.x - dpx, p0.y - dpy}, {p1.x - dpx, p1.y - dpy}, {p1.x + dpx, p1.y + dpy}};
        emia_cv_fill_convex4(W, H, pt, put);
    }
    emia_cv_plus1(W, H, x1, y1, put);
    emia_cv_plus1(W, H, x2, y2, put);
}
