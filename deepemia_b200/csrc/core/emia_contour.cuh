// emia_contour.cuh — external-contour extraction on a bit-packed crop (host/device).
//
// Replaces (reference call sites): cv2.findContours(mask, RETR_EXTERNAL, CHAIN_APPROX_SIMPLE) at
// src/functions/inference.py:1164-1167 (measurement loop) and :2605 (compactness pre-filter of
// deduplicate_masks_smart); cv2.contourArea at :1175 / src/utils/measurements.py:134 and cv2.arcLength at
// inference.py:2607 / measurements.py:135,163.
//
// Algorithm = Suzuki-Abe border following as OpenCV implements it (raster scan; an outer border starts at an
// unmarked foreground pixel whose left neighbour is background; in RETR_EXTERNAL mode the start is accepted
// only if the nearest already-marked pixel to its left on the same row is a "right-exit" pixel (negative
// label) or there is none; the border is followed with the 8-neighbour clockwise/counter-clockwise search and
// every visited pixel is labelled +2, or -126 when its east neighbour was examined and found empty).
// CHAIN_APPROX_SIMPLE keeps a point whenever the chain direction changes.
// State per pixel is held in two extra bit planes (marked / negative) instead of OpenCV's int8 image, so
// the scan for start pixels is word-parallel.
#pragma once
#include "emia_common.cuh"

#define EMIA_PACK_PT(x, y) ((uint32_t)(x) | ((uint32_t)(y) << 16))
#define EMIA_PT_X(p) ((int)((p) & 0xFFFFu))
#define EMIA_PT_Y(p) ((int)((p) >> 16))

struct EmiaContourOut {
    uint32_t* pts;     // packed vertices (frame coordinates), contours stored back-to-back in DISCOVERY order
    int cap_pts;
    int* cstart;       // cstart[k]..cstart[k+1] = vertex range of k-th discovered contour; size cap_contours+1
    int cap_contours;
    int n_contours;
    int n_pts;
    int overflow;      // set when a capacity was exceeded (results invalid)
    int max_len;       // longest contour (vertices)
    int store;         // 0: count only (pts / cstart untouched, capacities ignored)
};

EMIA_HD void emia_contour_emit(EmiaContourOut& o, int fx, int fy) {
    if (o.store) {
        if (o.n_pts < o.cap_pts) o.pts[o.n_pts] = EMIA_PACK_PT(fx, fy);
        else o.overflow = 1;
    }
    o.n_pts++;
}

struct EmiaMarks {
    uint32_t* mk;   // marked plane   (h * wwords words)
    uint32_t* ng;   // negative plane (h * wwords words)
    int wwords;
};
EMIA_HD int emia_marked(const EmiaMarks& m, int lx, int ly) {
    return (m.mk[ly * m.wwords + (lx >> 5)] >> (lx & 31)) & 1u;
}
EMIA_HD void emia_mark(const EmiaMarks& m, int lx, int ly, int negative) {
    const int w = ly * m.wwords + (lx >> 5);
    const uint32_t b = 1u << (lx & 31);
    m.mk[w] |= b;
    if (negative) m.ng[w] |= b;
}

// 8-neighbourhood, direction s: 0=E 1=NE 2=N 3=NW 4=W 5=SW 6=S 7=SE (y grows downwards); (delta + 1) packed 4 bits each
#define EMIA_DX(s) ((int)((0x21000122u >> (4 * (s))) & 0xFu) - 1)
#define EMIA_DY(s) ((int)((0x22210001u >> (4 * (s))) & 0xFu) - 1)

// bits of row ly at columns lx-1, lx, lx+1 (bit 0, 1, 2); anything outside the crop reads as 0
EMIA_HD uint32_t emia_row3(const EmiaBitView& v, int lx, int ly) {
    if ((unsigned)ly >= (unsigned)v.h) return 0u;
    const uint32_t* row = v.bits + (size_t)ly * v.pitch_words;
    if (lx == 0) return (row[0] << 1) & 7u;
    const int c = (lx - 1) >> 5, sh = (lx - 1) & 31;
    const uint32_t w0 = row[c];
    const uint32_t w1 = (c + 1 < v.wwords) ? row[c + 1] : 0u;
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(w0, w1, sh) & 7u;
#else
    return (uint32_t)(((((uint64_t)w1) << 32) | w0) >> sh) & 7u;
#endif
}
// 8-bit neighbour mask of local pixel (lx, ly): bit s set <=> the neighbour in direction s is foreground
EMIA_HD uint32_t emia_nbr8(const EmiaBitView& v, int lx, int ly) {
    const uint32_t up = emia_row3(v, lx, ly - 1), mid = emia_row3(v, lx, ly), dn = emia_row3(v, lx, ly + 1);
    return ((mid >> 2) & 1u) | (((up >> 2) & 1u) << 1) | (((up >> 1) & 1u) << 2) | ((up & 1u) << 3) | ((mid & 1u) << 4) |
           ((dn & 1u) << 5) | (((dn >> 1) & 1u) << 6) | (((dn >> 2) & 1u) << 7);
}

// All external contours of the crop, in discovery (raster) order.  OpenCV returns them in REVERSE discovery order;
// consumers iterate k = n_contours-1 .. 0.  mk/ng must hold h*wwords words each (zeroed here).
//
// Written as ONE flat loop over a two-phase state machine (scan for the next start pixel / one border-following step)
// so that the 32 instances of a warp share the instruction stream: a lane in the "follow" phase executes the same
// branch-free step (3x3 neighbour mask -> rotate -> count-trailing-zeros picks the next direction) whatever the shape.
// marks_zeroed != 0: the caller cleared both planes already (the kernels clear the whole buffer with coalesced stores).
EMIA_HD_NOINLINE void emia_find_external_contours(const EmiaBitView& v, uint32_t* mk, uint32_t* ng, EmiaContourOut& o, int marks_zeroed = 0) {
    const int ww = v.wwords;
    const int nw = v.h * ww;
    if (!marks_zeroed) for (int i = 0; i < nw; ++i) { mk[i] = 0u; ng[i] = 0u; }
    o.n_contours = 0; o.n_pts = 0; o.overflow = 0; o.max_len = 0;
    if (o.store) o.cstart[0] = 0;
    int y = 0, c = 0;
    uint32_t done_mask = 0u;            // bits of the current word already examined
    int phase = 0;                      // 0: scan, 1: follow
    int x0 = 0, y0 = 0, x1 = 0, y1 = 0, x3 = 0, y3 = 0, s = 0, prev_s = 0, before = 0;
    for (;;) {
        if (phase == 0) {
            if (y >= v.h) break;
            const uint32_t* row = v.bits + (size_t)y * v.pitch_words;
            const uint32_t F = row[c];
            const uint32_t left = (F << 1) | (c > 0 ? (row[c - 1] >> 31) : 0u);
            const uint32_t cand = F & ~left & ~mk[y * ww + c] & ~done_mask;
            if (cand == 0u) {
                done_mask = 0u;
                if (++c >= ww) { c = 0; ++y; }
                continue;
            }
            const int b = emia_ctz(cand);
            done_mask |= (b == 31) ? 0xFFFFFFFFu : ((2u << b) - 1u);
            const int x = c * 32 + b;
            // RETR_EXTERNAL: nearest marked pixel strictly left of x on this row must carry a negative label (or not exist)
            {
                int cc = c;
                uint32_t mw = mk[y * ww + cc] & ((b == 0) ? 0u : ((1u << b) - 1u));
                while (mw == 0u && cc > 0) { --cc; mw = mk[y * ww + cc]; }
                if (mw != 0u) {
                    const int hb = emia_msb(mw);
                    if (!((ng[y * ww + cc] >> hb) & 1u)) continue;   // inside an already-followed outer border
                }
            }
            if (o.store && o.n_contours >= o.cap_contours) { o.overflow = 1; return; }
            before = o.n_pts;
            const uint32_t N8 = emia_nbr8(v, x, y);
            // first neighbour clockwise from W: directions 3,2,1,0,7,6,5 -> bit k of M
            const uint32_t M = ((N8 >> 3) & 1u) | (((N8 >> 2) & 1u) << 1) | (((N8 >> 1) & 1u) << 2) | ((N8 & 1u) << 3) |
                               (((N8 >> 7) & 1u) << 4) | (((N8 >> 6) & 1u) << 5) | (((N8 >> 5) & 1u) << 6);
            const int wi = y * ww + (x >> 5);
            const uint32_t bit = 1u << (x & 31);
            if (M == 0u) {   // isolated pixel
                mk[wi] |= bit; ng[wi] |= bit;
                emia_contour_emit(o, v.x_origin + x, v.y_origin + y);
                o.n_contours++;
                if (o.n_pts - before > o.max_len) o.max_len = o.n_pts - before;
                if (o.store) o.cstart[o.n_contours] = o.n_pts;
                if (o.overflow) return;
                continue;
            }
            s = (3 - emia_ctz(M)) & 7;
            x0 = x; y0 = y; x1 = x + EMIA_DX(s); y1 = y + EMIA_DY(s);
            x3 = x0; y3 = y0;
            prev_s = s ^ 4;
            phase = 1;
        } else {
            const uint32_t N8 = emia_nbr8(v, x3, y3);
            const int s_end = s;
            const int r = (s_end + 1) & 7;
            const uint32_t rot = ((N8 >> r) | (N8 << (8 - r))) & 0xFFu;
            s = (s_end + 1 + emia_ctz(rot)) & 7;                 // next foreground neighbour counter-clockwise
            const int wi = y3 * ww + (x3 >> 5);
            const uint32_t bit = 1u << (x3 & 31);
            mk[wi] |= bit;                                        // label +2 ...
            if ((unsigned)(s - 1) < (unsigned)s_end) ng[wi] |= bit;   // ... or -126 when the east neighbour was examined empty
            if (s != prev_s) {
                emia_contour_emit(o, v.x_origin + x3, v.y_origin + y3);
                prev_s = s;
            }
            const int x4 = x3 + EMIA_DX(s), y4 = y3 + EMIA_DY(s);
            if (x4 == x0 && y4 == y0 && x3 == x1 && y3 == y1) {
                o.n_contours++;
                if (o.n_pts - before > o.max_len) o.max_len = o.n_pts - before;
                if (o.store) o.cstart[o.n_contours] = o.n_pts;
                if (o.overflow) return;
                phase = 0;
                continue;
            }
            x3 = x4; y3 = y4;
            s = (s + 4) & 7;
        }
    }
}

// cv2.contourArea on integer vertices: |shoelace| / 2 (exact; OpenCV accumulates integer-valued doubles).
EMIA_HD double emia_contour_area(const uint32_t* pts, int n) {
    if (n == 0) return 0.0;
    long long a = 0;
    int px = EMIA_PT_X(pts[n - 1]), py = EMIA_PT_Y(pts[n - 1]);
    for (int i = 0; i < n; ++i) {
        const int x = EMIA_PT_X(pts[i]), y = EMIA_PT_Y(pts[i]);
        a += (long long)px * y - (long long)py * x;
        px = x; py = y;
    }
    if (a < 0) a = -a;
    return (double)a * 0.5;
}

// cv2.arcLength(c, closed=True): per-segment float32 sqrt of float32 (dx*dx+dy*dy), accumulated in double,
// starting with the closing segment (last -> first).
// The segments of a CHAIN_APPROX_SIMPLE contour are runs in one of the 8 chain directions, so almost every term is
// either an axial run (sqrtf(k*k) == k exactly for k < 4096) or a diagonal run (fl32(sqrt(2 k^2)), taken from `diag`
// when the caller provides the table: diag[k] = sqrtf((float)(2*k*k)), k < EMIA_DIAG_TABLE); anything else goes through
// sqrtf, so the result is bit-identical to the plain loop for ANY vertex list.
#define EMIA_DIAG_TABLE 128
EMIA_HD double emia_arc_length_closed(const uint32_t* pts, int n, const float* diag = nullptr) {
    if (n <= 1) return 0.0;
    double per = 0.0;
    int px = EMIA_PT_X(pts[n - 1]), py = EMIA_PT_Y(pts[n - 1]);
    for (int i = 0; i < n; ++i) {
        const int x = EMIA_PT_X(pts[i]), y = EMIA_PT_Y(pts[i]);
        int dx = x - px, dy = y - py;
        dx = dx < 0 ? -dx : dx; dy = dy < 0 ? -dy : dy;
        float seg;
        if ((dx == 0 || dy == 0) && (dx | dy) < 4096) seg = (float)(dx | dy);
        else if (diag && dx == dy && dx < EMIA_DIAG_TABLE) seg = diag[dx];
        else {
            const float fx = (float)dx, fy = (float)dy;
            const float dx2 = fx * fx;
            const float dy2 = fy * fy;
            seg = sqrtf(dx2 + dy2);
        }
        per += (double)seg;
        px = x; py = y;
    }
    return per;
}
