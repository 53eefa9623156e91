"""Oracle: mask-IoU de-duplication (numpy restatement; test infrastructure only).

Follows src/functions/inference.py of the reference:
  * ``iou``                         :422-435
  * greedy in-order de-dup          :1453-1459 (``run_class_specific_inference`` tail)
  * ``deduplicate_masks_smart``     :2552-2677   (quirks Q1, Q2, Q10 of SURVEY.md Appendix A reproduced on purpose)
  * ``bboxes_overlap``              :2680-2694
  * ``calculate_iou``               :2697-2719
  * ``get_mask_bbox``               :2722-2733
"""
import cv2
import numpy as np


def iou(mask1, mask2):
    # inference.py:422-435 — full-frame logical and/or, 0 when the union is empty
    inter = np.logical_and(mask1, mask2).sum()
    union = np.logical_or(mask1, mask2).sum()
    return inter / union if union > 0 else 0


def greedy_inorder_dedup(masks, scores, target_class, iou_threshold):
    # inference.py:1453-1459 — keep a mask unless it overlaps an already kept one by more than the threshold
    kept, kept_scores, kept_classes, kept_idx = [], [], [], []
    for i, m in enumerate(masks):
        if not any(iou(m, u) > iou_threshold for u in kept):
            kept.append(m)
            kept_scores.append(scores[i])
            kept_classes.append(target_class)
            kept_idx.append(i)
    return kept, kept_scores, kept_classes, kept_idx


def get_mask_bbox(mask):
    # inference.py:2722-2733 / spatial_constraints.py:70-89 -> (y_min, x_min, y_max, x_max) or None
    rows = np.any(mask, axis=1)
    cols = np.any(mask, axis=0)
    if not rows.any() or not cols.any():
        return None
    ys = np.where(rows)[0]
    xs = np.where(cols)[0]
    return (ys[0], xs[0], ys[-1], xs[-1])


def bboxes_overlap(b1, b2):
    # inference.py:2680-2694: tuples are UNPACKED as (y_min, x_min, y_max, x_max)
    if b1 is None or b2 is None:
        return False
    y1a, x1a, y1b, x1b = b1
    y2a, x2a, y2b, x2b = b2
    if x1b < x2a or x2b < x1a:
        return False
    if y1b < y2a or y2b < y1a:
        return False
    return True


def calculate_iou(m1, m2, b1=None, b2=None):
    # inference.py:2697-2719
    if b1 is None:
        b1 = get_mask_bbox(m1)
    if b2 is None:
        b2 = get_mask_bbox(m2)
    if not bboxes_overlap(b1, b2):
        return 0.0
    inter = np.count_nonzero(m1 & m2)
    if inter == 0:
        return 0.0
    union = np.count_nonzero(m1 | m2)
    if union == 0:
        return 0.0
    return inter / union


def smart_prefilter_indices(masks, max_aspect_ratio=None):
    """Step 1 of deduplicate_masks_smart (inference.py:2573-2616): indices surviving the artifact pre-filter."""
    out = []
    for idx, mask in enumerate(masks):
        rows = np.any(mask, axis=1)
        cols = np.any(mask, axis=0)
        if not rows.any() or not cols.any():
            continue
        ys = np.where(rows)[0]
        xs = np.where(cols)[0]
        bw = xs[-1] - xs[0] + 1
        bh = ys[-1] - ys[0] + 1
        if bh == 0 or bw == 0:
            continue
        aspect = max(bw, bh) / min(bw, bh)
        if max_aspect_ratio and aspect > max_aspect_ratio:
            continue
        mask_area = np.sum(mask)
        contours, _ = cv2.findContours(np.asarray(mask).astype(np.uint8), cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
        if len(contours) > 0:
            perimeter = cv2.arcLength(contours[0], True)     # Q10: first RETURNED contour only
            if perimeter > 0:
                compactness = (4 * np.pi * mask_area) / (perimeter ** 2)
                if compactness < 0.15:
                    continue
        out.append(idx)
    return out


def deduplicate_masks_smart(masks, scores, classes, iou_threshold=0.4, max_aspect_ratio=None, return_indices=False):
    """inference.py:2552-2677.  Returns (masks, scores, classes) in keep order [+ original indices]."""
    if len(masks) == 0:
        return ([], [], [], []) if return_indices else ([], [], [])
    surv = smart_prefilter_indices(masks, max_aspect_ratio)
    masks_f = [masks[i] for i in surv]
    scores_f = [scores[i] for i in surv]
    classes_f = [classes[i] for i in surv]
    if len(masks_f) == 0:
        return ([], [], [], []) if return_indices else ([], [], [])
    # Q1: boxes are STORED as (y_min, y_max, x_min, x_max) but consumed by bboxes_overlap as (y_min, x_min, y_max, x_max)
    boxes = []
    for m in masks_f:
        rows = np.any(m, axis=1)
        cols = np.any(m, axis=0)
        if rows.any() and cols.any():
            ys = np.where(rows)[0]
            xs = np.where(cols)[0]
            boxes.append((ys[0], ys[-1], xs[0], xs[-1]))
        else:
            boxes.append(None)
    order = np.argsort(scores_f)[::-1]
    keep, removed = [], set()
    for idx in order:
        if idx in removed:
            continue
        keep.append(idx)
        # Q2: the slice start is the VALUE of idx (an index into the filtered list), not its position in `order`
        for other in order[idx + 1:]:
            if other in removed:
                continue
            if classes_f[other] != classes_f[idx]:
                continue
            if not bboxes_overlap(boxes[idx], boxes[other]):
                continue
            if calculate_iou(masks_f[idx], masks_f[other], boxes[idx], boxes[other]) > iou_threshold:
                removed.add(other)
    res = ([masks_f[i] for i in keep], [scores_f[i] for i in keep], [classes_f[i] for i in keep])
    if return_indices:
        return res + ([surv[i] for i in keep],)
    return res
