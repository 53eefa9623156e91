// emia_paste.cuh — Detectron2-style full-resolution mask paste + threshold, per pixel (host/device).
//
// Replaces (reference): the third-party step reached through `predictor(image)` at
// src/functions/inference.py:1395,1398,1507,1669,2107 — Detectron2 0.6 `detector_postprocess`
// (box scale + clip + non-empty filter) followed by `paste_masks_in_image(masks[:,0], boxes, (H,W), 0.5)`
// (`_do_paste_mask`: grid = ((p + 0.5) - x0) / (x1 - x0) * 2 - 1, `F.grid_sample(bilinear, zeros,
// align_corners=False)`, `>= 0.5`).  SURVEY.md Appendix B.1.
//
// Bit-exactness: the oracle is PyTorch's CPU grid_sample; its float32 operation sequence is
//   g  = ((p + 0.5) - x0) / (x1 - x0) * 2 - 1            (separate torch ops: each individually rounded)
//   ix = fma(g + 1, size/2, -0.5)
//   x_w = floor(ix); tw = ix - x_w; te = 1 - tw           (same for y: tn, ts)
//   nw = ts*te, ne = ts*tw, sw = tn*te, se = tn*tw        (each rounded)
//   out = fma(v_se, se, fma(v_sw, sw, fma(v_ne, ne, v_nw * nw)))      taps outside [0,size) read as 0
// The library is built with -fmad=false, so only the explicit emia_fmaf calls fuse.
#pragma once
#include "emia_common.cuh"

#define EMIA_MASK_SIDE 28

struct EmiaPasteBox {
    float x0, y0, x1, y1;      // scaled + clipped box
    int rx0, ry0, rx1, ry1;    // integer sampling region [rx0, rx1) x [ry0, ry1)
    int valid;                 // Boxes.nonempty()
};

EMIA_HD float emia_clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }

// detector_postprocess box handling + the CPU skip_empty region of _do_paste_mask
EMIA_HD EmiaPasteBox emia_paste_prepare(float bx0, float by0, float bx1, float by1, float scale_x, float scale_y, int W, int H) {
    EmiaPasteBox b;
    b.x0 = emia_clampf(bx0 * scale_x, 0.f, (float)W);
    b.x1 = emia_clampf(bx1 * scale_x, 0.f, (float)W);
    b.y0 = emia_clampf(by0 * scale_y, 0.f, (float)H);
    b.y1 = emia_clampf(by1 * scale_y, 0.f, (float)H);
    b.valid = ((b.x1 - b.x0) > 0.f) && ((b.y1 - b.y0) > 0.f);
    b.rx0 = emia_max((int)floorf(b.x0) - 1, 0);
    b.ry0 = emia_max((int)floorf(b.y0) - 1, 0);
    b.rx1 = emia_min((int)ceilf(b.x1) + 1, W);
    b.ry1 = emia_min((int)ceilf(b.y1) + 1, H);
    if (!b.valid) { b.rx0 = b.ry0 = b.rx1 = b.ry1 = 0; }
    return b;
}

struct EmiaAxisTap {
    int i0;        // floor(ix)
    float w1;      // ix - i0   (weight of tap i0+1)
    float w0;      // 1 - w1    (weight of tap i0)
};

// one axis of the sampling grid for integer pixel p
EMIA_HD EmiaAxisTap emia_paste_axis(int p, float lo, float hi) {
    const float c = (float)p + 0.5f;
    const float t = (c - lo) / (hi - lo);
    const float g = t * 2.f - 1.f;
    const float ix = emia_fmaf(g + 1.f, (float)EMIA_MASK_SIDE * 0.5f, -0.5f);
    const float fl = floorf(ix);
    EmiaAxisTap a;
    a.i0 = (int)fl;
    a.w1 = ix - fl;
    a.w0 = 1.f - a.w1;
    return a;
}

EMIA_HD float emia_paste_tap(const float* prob, int x, int y) {
    return ((unsigned)x < (unsigned)EMIA_MASK_SIDE && (unsigned)y < (unsigned)EMIA_MASK_SIDE) ? prob[y * EMIA_MASK_SIDE + x] : 0.f;
}

// interpolated probability at (column tap ax, row tap ay); prob = 28x28 row-major
EMIA_HD float emia_paste_sample(const float* prob, const EmiaAxisTap& ax, const EmiaAxisTap& ay) {
    const float nw = ay.w0 * ax.w0;
    const float ne = ay.w0 * ax.w1;
    const float sw = ay.w1 * ax.w0;
    const float se = ay.w1 * ax.w1;
    const float v_nw = emia_paste_tap(prob, ax.i0, ay.i0);
    const float v_ne = emia_paste_tap(prob, ax.i0 + 1, ay.i0);
    const float v_sw = emia_paste_tap(prob, ax.i0, ay.i0 + 1);
    const float v_se = emia_paste_tap(prob, ax.i0 + 1, ay.i0 + 1);
    float acc = v_nw * nw;
    acc = emia_fmaf(v_ne, ne, acc);
    acc = emia_fmaf(v_sw, sw, acc);
    acc = emia_fmaf(v_se, se, acc);
    return acc;
}
