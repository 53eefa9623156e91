// emia_nms.cuh — pair primitives for mask-IoU de-duplication / spatial constraints (host/device).
//
// Replaces (reference): iou (src/functions/inference.py:422-435), calculate_iou (:2697-2719,
// src/utils/spatial_constraints.py:118-153), calculate_containment (spatial_constraints.py:156-189),
// bboxes_overlap (inference.py:2680-2694 with the Q1 tuple-order quirk; spatial_constraints.py:92-115 without).
#pragma once
#include "emia_common.cuh"

struct EmiaCropRef {
    const uint32_t* w;   // ch * cw words
    int ry0, wc0, ch, cw;
};

// popcount(a AND b): crops are word-aligned with the frame, so no shifting is needed
EMIA_HD int emia_crop_inter(const EmiaCropRef& a, const EmiaCropRef& b) {
    const int r0 = emia_max(a.ry0, b.ry0), r1 = emia_min(a.ry0 + a.ch, b.ry0 + b.ch);
    const int c0 = emia_max(a.wc0, b.wc0), c1 = emia_min(a.wc0 + a.cw, b.wc0 + b.cw);
    if (r0 >= r1 || c0 >= c1) return 0;
    int cnt = 0;
    for (int r = r0; r < r1; ++r) {
        const uint32_t* pa = a.w + (size_t)(r - a.ry0) * a.cw - a.wc0;
        const uint32_t* pb = b.w + (size_t)(r - b.ry0) * b.cw - b.wc0;
        for (int c = c0; c < c1; ++c) cnt += emia_popc(pa[c] & pb[c]);
    }
    return cnt;
}

// boxes are (y_min, x_min, y_max, x_max), inclusive; y_min < 0 encodes "None" (empty mask)
EMIA_HD bool emia_bbox_overlap(const int* A, const int* B) {
    if (A[0] < 0 || B[0] < 0) return false;
    if (A[3] < B[1] || B[3] < A[1]) return false;
    if (A[2] < B[0] || B[2] < A[0]) return false;
    return true;
}
// Q1: deduplicate_masks_smart stores (y_min, y_max, x_min, x_max) and bboxes_overlap unpacks the tuple as
// (y_min, x_min, y_max, x_max): the test actually evaluated mixes x and y extents.
EMIA_HD bool emia_bbox_overlap_q1(const int* A, const int* B) {
    if (A[0] < 0 || B[0] < 0) return false;
    const int a_ymin = A[0], a_xmin = A[1], a_ymax = A[2], a_xmax = A[3];
    const int b_ymin = B[0], b_xmin = B[1], b_ymax = B[2], b_xmax = B[3];
    if (a_xmax < b_ymax || b_xmax < a_ymax) return false;
    if (a_xmin < b_ymin || b_xmin < a_ymin) return false;
    return true;
}
