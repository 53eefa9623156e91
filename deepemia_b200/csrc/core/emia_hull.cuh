// emia_hull.cuh — convex hull, rotating-calipers min-area rectangle, box corners (host/device).
//
// Replaces (reference call sites): cv2.minAreaRect / cv2.boxPoints at src/utils/measurements.py:138-139,
// followed by the int truncation (:140), imutils.perspective.order_points (:141) and the midpoint /
// Euclidean-distance arithmetic (:142-148) that yields dA, dB.
//
// The reference truncates the float32 box corners to integers, which makes Length / Width / Feret / Aspect /
// Roundness discontinuous in the corner values; the float32 operation order of OpenCV's implementation is
// therefore followed step by step (Sklansky hull over points sorted by (x, y, index); calipers started at the
// hull's extreme points; `area <= minarea` tie rule; corner reconstruction from (center, size, angle)).
#pragma once
#include "emia_common.cuh"
#include "emia_contour.cuh"

// ---- sort point indices by (x, y, index): heap sort on 64-bit keys (total order => algorithm independent)
EMIA_HD void emia_sort_keys(uint64_t* a, int n) {
    for (int start = n / 2 - 1; start >= 0; --start) {
        int root = start;
        for (;;) {
            int child = 2 * root + 1;
            if (child >= n) break;
            if (child + 1 < n && a[child] < a[child + 1]) child++;
            if (a[root] < a[child]) { uint64_t t = a[root]; a[root] = a[child]; a[child] = t; root = child; }
            else break;
        }
    }
    for (int end = n - 1; end > 0; --end) {
        uint64_t t = a[0]; a[0] = a[end]; a[end] = t;
        int root = 0;
        for (;;) {
            int child = 2 * root + 1;
            if (child >= end) break;
            if (child + 1 < end && a[child] < a[child + 1]) child++;
            if (a[root] < a[child]) { uint64_t t2 = a[root]; a[root] = a[child]; a[child] = t2; root = child; }
            else break;
        }
    }
}
#define EMIA_KEY(x, y, i) (((uint64_t)(uint32_t)(x) << 44) | ((uint64_t)(uint32_t)(y) << 24) | (uint64_t)(uint32_t)(i))
#define EMIA_KEY_X(k) ((int)((k) >> 44))
#define EMIA_KEY_Y(k) ((int)(((k) >> 24) & 0xFFFFF))
#define EMIA_KEY_I(k) ((int)((k) & 0xFFFFFF))

EMIA_HD int emia_sign_ll(long long v) { return (v > 0) - (v < 0); }
// Compact 32-bit keys for contours whose extent is below 4096 x 4096 and that have at most 256 vertices (coordinates relative to
// the contour's own minimum): (x << 20) | (y << 8) | index.  Same order, same differences, half the shared-memory traffic and
// 32-bit arithmetic in the chains.  The accessors below are overloaded on the key type.
#define EMIA_KEY32(x, y, i) (((uint32_t)(x) << 20) | ((uint32_t)(y) << 8) | (uint32_t)(i))
EMIA_HD int emia_kx(uint64_t k) { return EMIA_KEY_X(k); }
EMIA_HD int emia_ky(uint64_t k) { return EMIA_KEY_Y(k); }
EMIA_HD int emia_ki(uint64_t k) { return EMIA_KEY_I(k); }
EMIA_HD int emia_kx(uint32_t k) { return (int)(k >> 20); }
EMIA_HD int emia_ky(uint32_t k) { return (int)((k >> 8) & 0xFFFu); }
EMIA_HD int emia_ki(uint32_t k) { return (int)(k & 0xFFu); }
EMIA_HD int emia_convexity_sign(uint64_t, int ay, int bx, int ax, int by) { return emia_sign_ll((long long)ay * bx - (long long)ax * by); }
EMIA_HD int emia_convexity_sign(uint32_t, int ay, int bx, int ax, int by) { const int v = ay * bx - ax * by; return (v > 0) - (v < 0); }

// One monotone chain of the hull (OpenCV's Sklansky_ on the sorted array).  Returns stack size.
template <typename S, typename K>
EMIA_HD int emia_sklansky(const K* arr, int start, int end, S* stack, int nsign, int sign2) {
    const int incr = end > start ? 1 : -1;
    int pprev = start, pcur = pprev + incr, pnext = pcur + incr;
    int stacksize = 3;
    if (start == end || (emia_kx(arr[start]) == emia_kx(arr[end]) && emia_ky(arr[start]) == emia_ky(arr[end]))) {
        stack[0] = start;
        return 1;
    }
    stack[0] = pprev; stack[1] = pcur; stack[2] = pnext;
    end += incr;
    while (pnext != end) {
        const int cury = emia_ky(arr[pcur]);
        const int nexty = emia_ky(arr[pnext]);
        const int by = nexty - cury;
        if (((by > 0) - (by < 0)) != nsign) {
            const int ax = emia_kx(arr[pcur]) - emia_kx(arr[pprev]);
            const int bx = emia_kx(arr[pnext]) - emia_kx(arr[pcur]);
            const int ay = cury - emia_ky(arr[pprev]);
            if (emia_convexity_sign(arr[pcur], ay, bx, ax, by) == sign2 && (ax != 0 || ay != 0)) {
                pprev = pcur;
                pcur = pnext;
                pnext += incr;
                stack[stacksize] = pnext;
                stacksize++;
            } else {
                if (pprev == start) {
                    pcur = pnext;
                    stack[1] = pcur;
                    pnext += incr;
                    stack[2] = pnext;
                } else {
                    stack[stacksize - 2] = pnext;
                    pcur = pprev;
                    pprev = stack[stacksize - 4];
                    stacksize--;
                }
            }
        } else {
            pnext += incr;
            stack[stacksize - 1] = pnext;
        }
    }
    return --stacksize;
}

// The hull is assembled from four monotone chains (tl: leftmost -> topmost, tr: rightmost -> topmost, bl / br likewise for
// the bottom).  The three helpers below are OpenCV's assembly steps; they are shared by the serial emia_convex_hull and
// the warp-cooperative hull kernel (which runs the four chains on four lanes, keys and stacks in shared memory).
template <typename S, typename K, typename H>
EMIA_HD int emia_hull_emit_upper(const K* keys, int clockwise, S* tl_stack, int tl_count, S* tr_stack, int tr_count,
                                 H* hull, int* nout_io) {
    int nout = *nout_io;
    if (!clockwise) {
        S* ts = tl_stack; tl_stack = tr_stack; tr_stack = ts;
        int tc = tl_count; tl_count = tr_count; tr_count = tc;
    }
    for (int i = 0; i < tl_count - 1; ++i) hull[nout++] = (H)emia_ki(keys[tl_stack[i]]);
    for (int i = tr_count - 1; i > 0; --i) hull[nout++] = (H)emia_ki(keys[tr_stack[i]]);
    *nout_io = nout;
    return tr_count > 2 ? (int)tr_stack[1] : tl_count > 2 ? (int)tl_stack[tl_count - 2] : -1;   // stop_idx
}
template <typename S, typename K, typename H>
EMIA_HD void emia_hull_emit_lower(const K* keys, int clockwise, S* bl_stack, int bl_count, S* br_stack, int br_count,
                                  int stop_idx, H* hull, int* nout_io) {
    int nout = *nout_io;
    if (clockwise) {
        S* ts = bl_stack; bl_stack = br_stack; br_stack = ts;
        int tc = bl_count; bl_count = br_count; br_count = tc;
    }
    if (stop_idx >= 0) {
        const int check_idx = bl_count > 2 ? (int)bl_stack[1] : bl_count + br_count > 2 ? (int)br_stack[2 - bl_count] : -1;
        if (check_idx == stop_idx ||
            (check_idx >= 0 && emia_kx(keys[check_idx]) == emia_kx(keys[stop_idx]) &&
             emia_ky(keys[check_idx]) == emia_ky(keys[stop_idx]))) {
            bl_count = emia_min(bl_count, 2);
            br_count = emia_min(br_count, 2);
        }
    }
    for (int i = 0; i < bl_count - 1; ++i) hull[nout++] = (H)emia_ki(keys[bl_stack[i]]);
    for (int i = br_count - 1; i > 0; --i) hull[nout++] = (H)emia_ki(keys[br_stack[i]]);
    *nout_io = nout;
}
// cyclic shift so that indices ascend/descend when possible (H: int, or uint8_t for the packed hull kernel)
template <typename H>
EMIA_HD void emia_hull_cyclic_shift(H* hull, int nout, H* tmp) {
    if (nout < 3) return;
    int min_idx = 0, max_idx = 0, lt = 0;
    for (int i = 1; i < nout; ++i) {
        const int idx = hull[i];
        lt += (int)hull[i - 1] < idx;
        if (lt > 1 && lt <= i - 2) break;
        if (idx < hull[min_idx]) min_idx = i;
        if (idx > hull[max_idx]) max_idx = i;
    }
    int mmdist = max_idx - min_idx; if (mmdist < 0) mmdist = -mmdist;
    if ((mmdist == 1 || mmdist == nout - 1) && (lt <= 1 || lt >= nout - 2)) {
        const int ascending = (max_idx + 1) % nout == min_idx;
        const int i0 = ascending ? min_idx : max_idx;
        int j = i0;
        if (i0 > 0) {
            int i;
            for (i = 0; i < nout; ++i) {
                const int curr_idx = tmp[i] = hull[j];
                const int next_j = j + 1 < nout ? j + 1 : 0;
                const int next_idx = hull[next_j];
                if (i < nout - 1 && (ascending != (curr_idx < next_idx))) break;
                j = next_j;
            }
            if (i == nout) for (int k = 0; k < nout; ++k) hull[k] = tmp[k];
        }
    }
}

// Convex hull of n packed integer points (counter-clockwise flag as cv2.convexHull(clockwise=...)).
// keys: scratch n; stack: scratch n+2; hull: out, original point indices (capacity n); tmp: scratch n.
// Returns hull size.
// presorted != 0: keys[] already holds the n keys in ascending order (sorted by a cooperative kernel beforehand).
EMIA_HD_NOINLINE int emia_convex_hull(const uint32_t* pts, int n, int clockwise, uint64_t* keys, int* stack, int* hull, int* tmp, int presorted = 0) {
    if (n == 0) return 0;
    if (!presorted) {
        for (int i = 0; i < n; ++i) keys[i] = EMIA_KEY(EMIA_PT_X(pts[i]), EMIA_PT_Y(pts[i]), i);
        emia_sort_keys(keys, n);
    }
    int miny_ind = 0, maxy_ind = 0;
    for (int i = 1; i < n; ++i) {
        const int y = EMIA_KEY_Y(keys[i]);
        if (EMIA_KEY_Y(keys[miny_ind]) > y) miny_ind = i;
        if (EMIA_KEY_Y(keys[maxy_ind]) < y) maxy_ind = i;
    }
    int nout = 0;
    if (EMIA_KEY_X(keys[0]) == EMIA_KEY_X(keys[n - 1]) && EMIA_KEY_Y(keys[0]) == EMIA_KEY_Y(keys[n - 1])) {
        hull[nout++] = 0;
        return nout;
    }
    // upper half
    int* tl_stack = stack;
    const int tl_count = emia_sklansky(keys, 0, maxy_ind, tl_stack, -1, 1);
    int* tr_stack = stack + tl_count;
    const int tr_count = emia_sklansky(keys, n - 1, maxy_ind, tr_stack, -1, -1);
    const int stop_idx = emia_hull_emit_upper(keys, clockwise, tl_stack, tl_count, tr_stack, tr_count, hull, &nout);
    // lower half.  NOTE: the upper-half stacks are dead from here on (their content was copied to hull).
    int* bl_stack = stack;
    const int bl_count = emia_sklansky(keys, 0, miny_ind, bl_stack, 1, -1);
    int* br_stack = stack + bl_count;
    const int br_count = emia_sklansky(keys, n - 1, miny_ind, br_stack, 1, 1);
    emia_hull_emit_lower(keys, clockwise, bl_stack, bl_count, br_stack, br_count, stop_idx, hull, &nout);
    emia_hull_cyclic_shift(hull, nout, tmp);
    return nout;
}

// ---- rotating calipers, min-area rectangle.  hq = hull points packed (x | y << 16) in hull order, n >= 3.
// out[6] = corner (x,y), vec1 (x,y), vec2 (x,y).  OpenCV keeps the edge vectors and their inverse lengths in two arrays; here
// they are recomputed where they are used (the same float32 / float64 operations on the same integer-valued inputs, so the
// results are bit-identical), which leaves the packed hull points as the only memory the loop touches — they fit in shared
// memory in the hull kernel and in a 4n-byte scratch region in the serial path.
struct EmiaHullEdge { float x, y; };
EMIA_HD EmiaHullEdge emia_hull_edge(const uint32_t* hq, int n, int i) {
    const int nx = (i + 1 < n) ? i + 1 : 0;
    EmiaHullEdge e;
    e.x = (float)(EMIA_PT_X(hq[nx]) - EMIA_PT_X(hq[i]));          // exact: |difference| < 2^16
    e.y = (float)(EMIA_PT_Y(hq[nx]) - EMIA_PT_Y(hq[i]));
    return e;
}
EMIA_HD_NOINLINE void emia_rotating_calipers(const uint32_t* hq, int n, float* out) {
    float minarea = FLT_MAX;
    int left = 0, bottom = 0, right = 0, top = 0;
    int seq[4] = {-1, -1, -1, -1};
    float orientation = 0.f;
    float base_a, base_b = 0.f;
    float left_x, right_x, top_y, bottom_y;
    left_x = right_x = (float)EMIA_PT_X(hq[0]);
    top_y = bottom_y = (float)EMIA_PT_Y(hq[0]);
    // saved best
    int best_left = 0, best_bottom = 0;
    float best_a = 0.f, best_w = 0.f, best_b = 0.f, best_h = 0.f;

    for (int i = 0; i < n; ++i) {
        const float p0x = (float)EMIA_PT_X(hq[i]), p0y = (float)EMIA_PT_Y(hq[i]);
        if (p0x < left_x) { left_x = p0x; left = i; }
        if (p0x > right_x) { right_x = p0x; right = i; }
        if (p0y > top_y) { top_y = p0y; top = i; }
        if (p0y < bottom_y) { bottom_y = p0y; bottom = i; }
    }
    {
        const EmiaHullEdge last = emia_hull_edge(hq, n, n - 1);
        double ax = last.x, ay = last.y;
        for (int i = 0; i < n; ++i) {
            const EmiaHullEdge e = emia_hull_edge(hq, n, i);
            const double bx = e.x, by = e.y;
            const double convexity = ax * by - ay * bx;
            if (convexity != 0) { orientation = (convexity > 0) ? 1.f : -1.f; break; }
            ax = bx; ay = by;
        }
    }
    base_a = orientation;
    seq[0] = bottom; seq[1] = right; seq[2] = top; seq[3] = left;

    for (int k = 0; k < n; ++k) {
        // Choose the calipers side that makes the smallest angle with its polygon edge: rotate the four edge
        // vectors into a common frame and compare them pairwise by the sign of a cross product (exact for
        // integer-valued hull points; no cosine/inverse-length rounding involved).
        const EmiaHullEdge e0 = emia_hull_edge(hq, n, seq[0]), e1 = emia_hull_edge(hq, n, seq[1]);
        const EmiaHullEdge e2 = emia_hull_edge(hq, n, seq[2]), e3 = emia_hull_edge(hq, n, seq[3]);
        float rvx[4], rvy[4];
        rvx[0] = e0.x;   rvy[0] = e0.y;
        rvx[1] = e1.y;   rvy[1] = -e1.x;      // rotated 90 deg clockwise
        rvx[2] = -e2.x;  rvy[2] = -e2.y;      // rotated 180 deg
        rvx[3] = -e3.y;  rvy[3] = e3.x;       // rotated 90 deg counter-clockwise
        int main_element = 0;
        float mx = rvx[0], my = rvy[0];
        for (int i = 1; i < 4; ++i) {
            const float tx = rvy[i], ty = -rvx[i];
            if (tx * mx + ty * my < 0) { main_element = i; mx = rvx[i]; my = rvy[i]; }
        }
        {
            const EmiaHullEdge lead = main_element == 0 ? e0 : main_element == 1 ? e1 : main_element == 2 ? e2 : e3;
            const double dx = (double)lead.x, dy = (double)lead.y;
            const float inv_len = (float)(1. / sqrt(dx * dx + dy * dy));
            const float lead_x = lead.x * inv_len;
            const float lead_y = lead.y * inv_len;
            switch (main_element) {
                case 0: base_a = lead_x; base_b = lead_y; break;
                case 1: base_a = lead_y; base_b = -lead_x; break;
                case 2: base_a = -lead_x; base_b = -lead_y; break;
                default: base_a = -lead_y; base_b = lead_x; break;
            }
        }
        seq[main_element] += 1;
        seq[main_element] = (seq[main_element] == n) ? 0 : seq[main_element];
        {
            float dx = (float)EMIA_PT_X(hq[seq[1]]) - (float)EMIA_PT_X(hq[seq[3]]);
            float dy = (float)EMIA_PT_Y(hq[seq[1]]) - (float)EMIA_PT_Y(hq[seq[3]]);
            const float width = dx * base_a + dy * base_b;
            dx = (float)EMIA_PT_X(hq[seq[2]]) - (float)EMIA_PT_X(hq[seq[0]]);
            dy = (float)EMIA_PT_Y(hq[seq[2]]) - (float)EMIA_PT_Y(hq[seq[0]]);
            const float height = -dx * base_b + dy * base_a;
            const float area = width * height;
            if (area <= minarea) {
                minarea = area;
                best_left = seq[3];
                best_a = base_a; best_w = width; best_b = base_b; best_h = height;
                best_bottom = seq[0];
            }
        }
    }
    {
        const float A1 = best_a, B1 = best_b;
        const float A2 = -best_b, B2 = best_a;
        const float C1 = A1 * (float)EMIA_PT_X(hq[best_left]) + (float)EMIA_PT_Y(hq[best_left]) * B1;
        const float C2 = A2 * (float)EMIA_PT_X(hq[best_bottom]) + (float)EMIA_PT_Y(hq[best_bottom]) * B2;
        const float idet = 1.f / (A1 * B2 - A2 * B1);
        const float px = (C1 * B2 - C2 * B1) * idet;
        const float py = (A1 * C2 - A2 * C1) * idet;
        out[0] = px; out[1] = py;
        out[2] = A1 * best_w; out[3] = B1 * best_w;
        out[4] = A2 * best_h; out[5] = B2 * best_h;
    }
}

struct EmiaRotRect { float cx, cy, w, h, angle; };

// cv2.minAreaRect (OpenCV 4.13) from the calipers result out[6] (n > 2) or the first two hull points hq01 (n <= 2).
// The angle is evaluated in double from the first side vector and normalised into [-90, 0) by quarter turns
// (each turn swaps width and height) before the single rounding to float32.
EMIA_HD EmiaRotRect emia_min_area_rect_finish(int n, const float* out, const uint32_t* hq01) {
    EmiaRotRect box; box.cx = box.cy = box.w = box.h = box.angle = 0.f;
    double ang = 0.0;
    if (n > 2) {
        box.cx = out[0] + (out[2] + out[4]) * 0.5f;
        box.cy = out[1] + (out[3] + out[5]) * 0.5f;
        box.w = (float)sqrt((double)out[2] * out[2] + (double)out[3] * out[3]);
        box.h = (float)sqrt((double)out[4] * out[4] + (double)out[5] * out[5]);
        ang = atan2((double)out[3], (double)out[2]);
    } else if (n == 2) {
        const float x0 = (float)EMIA_PT_X(hq01[0]), y0 = (float)EMIA_PT_Y(hq01[0]);
        const float x1 = (float)EMIA_PT_X(hq01[1]), y1 = (float)EMIA_PT_Y(hq01[1]);
        box.cx = (x0 + x1) * 0.5f;
        box.cy = (y0 + y1) * 0.5f;
        const double dx = x1 - x0;
        const double dy = y1 - y0;
        box.w = (float)sqrt(dx * dx + dy * dy);
        box.h = 0;
        ang = atan2(dy, dx);
    } else if (n == 1) {
        box.cx = (float)EMIA_PT_X(hq01[0]); box.cy = (float)EMIA_PT_Y(hq01[0]);
    }
    ang = ang * 180 / M_PI;
    while (ang >= 0.0) { ang -= 90.0; const float t = box.w; box.w = box.h; box.h = t; }
    while (ang < -90.0) { ang += 90.0; const float t = box.w; box.w = box.h; box.h = t; }
    box.angle = (float)ang;
    return box;
}
// ... on the counter-clockwise hull points hq (packed), n = hull size.
EMIA_HD EmiaRotRect emia_min_area_rect_from_hull(const uint32_t* hq, int n) {
    float out[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (n > 2) emia_rotating_calipers(hq, n, out);
    return emia_min_area_rect_finish(n, out, hq);
}

// cv2.boxPoints: 4 corners as float32 (x,y interleaved)
EMIA_HD void emia_box_points(const EmiaRotRect& r, float* pt) {
    const double ang = r.angle * M_PI / 180.;
    const float b = (float)cos(ang) * 0.5f;
    const float a = (float)sin(ang) * 0.5f;
    pt[0] = r.cx - a * r.h - b * r.w;
    pt[1] = r.cy + b * r.h - a * r.w;
    pt[2] = r.cx + a * r.h - b * r.w;
    pt[3] = r.cy - b * r.h - a * r.w;
    pt[4] = 2 * r.cx - pt[0];
    pt[5] = 2 * r.cy - pt[1];
    pt[6] = 2 * r.cx - pt[2];
    pt[7] = 2 * r.cy - pt[3];
}
