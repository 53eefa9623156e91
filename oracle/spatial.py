"""Oracle: class containment / overlap spatial-constraint filtering (numpy restatement; test infrastructure only).

Follows src/utils/spatial_constraints.py of the reference:
  * ``get_mask_bbox`` :70-89, ``bboxes_overlap`` :92-115 (consistent tuple order here), ``calculate_iou`` :118-153,
    ``calculate_containment`` :156-189
  * ``filter_by_overlap_rules`` :192-277, ``filter_by_containment_rules`` :280-398
  * ``apply_spatial_constraints`` :401-460 with the rule set passed explicitly instead of read from YAML (:21-67).
"""
import numpy as np

from .dedup import get_mask_bbox


def bboxes_overlap(b1, b2):
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
    return inter / union if union else 0.0


def calculate_containment(child, parent, cb=None, pb=None):
    if cb is None:
        cb = get_mask_bbox(child)
    if pb is None:
        pb = get_mask_bbox(parent)
    if not bboxes_overlap(cb, pb):
        return 0.0
    area = np.count_nonzero(child)
    if area == 0:
        return 0.0
    return np.count_nonzero(child & parent) / area


def filter_by_overlap_rules(masks, scores, classes, overlap_rules):
    """:192-277.  Returns the set of removed indices."""
    removed = set()
    if not overlap_rules:
        return removed
    boxes = [get_mask_bbox(m) for m in masks]
    groups = {}
    for i, c in enumerate(classes):
        groups.setdefault(c, []).append(i)
    for cls, idxs in groups.items():
        if cls not in overlap_rules:
            continue
        rule = overlap_rules[cls]
        allow = rule.get('allow_overlap', True)
        max_iou = rule.get('max_iou_threshold', 0.5)
        if allow and max_iou >= 0.9:
            continue
        order = sorted(idxs, key=lambda i: scores[i], reverse=True)   # stable
        for p, i1 in enumerate(order):
            if i1 in removed:
                continue
            for i2 in order[p + 1:]:
                if i2 in removed:
                    continue
                if not bboxes_overlap(boxes[i1], boxes[i2]):
                    continue
                if calculate_iou(masks[i1], masks[i2], boxes[i1], boxes[i2]) > max_iou:
                    removed.add(i2)
    return removed


def filter_by_containment_rules(masks, scores, classes, containment_rules, containment_threshold=0.95):
    """:280-398.  Returns the set of removed indices."""
    removed = set()
    if not containment_rules:
        return removed
    boxes = [get_mask_bbox(m) for m in masks]
    by_class = {}
    for i, c in enumerate(classes):
        by_class.setdefault(c, []).append(i)
    for child_cls, parent_cls in containment_rules.items():
        if child_cls not in by_class:
            continue
        if parent_cls not in by_class:
            removed.update(by_class[child_cls])
            continue
        parents = [(p, boxes[p]) for p in by_class[parent_cls] if p not in removed and boxes[p] is not None]
        for ch in by_class[child_cls]:
            if ch in removed:
                continue
            if boxes[ch] is None:
                removed.add(ch)
                continue
            best = 0.0
            for p, pb in parents:
                if p in removed:
                    continue
                if not bboxes_overlap(boxes[ch], pb):
                    continue
                c = calculate_containment(masks[ch], masks[p], boxes[ch], pb)
                if c > best:
                    best = c
            if best < containment_threshold:
                removed.add(ch)
    return removed


def apply_spatial_constraints(masks, scores, classes, rules):
    """:401-460 with ``rules`` = {'enabled', 'overlap_rules', 'containment_rules', 'containment_threshold'}.
    Returns (masks, scores, classes, kept_indices)."""
    idx = list(range(len(masks)))
    if not masks or not rules or not rules.get('enabled', False):
        return masks, scores, classes, idx
    thr = rules.get('containment_threshold', 0.95)
    orules = rules.get('overlap_rules', {})
    if orules:
        rem = filter_by_overlap_rules(masks, scores, classes, orules)
        keep = [i for i in range(len(masks)) if i not in rem]
        masks = [masks[i] for i in keep]; scores = [scores[i] for i in keep]
        classes = [classes[i] for i in keep]; idx = [idx[i] for i in keep]
    crules = rules.get('containment_rules', {})
    if crules:
        rem = filter_by_containment_rules(masks, scores, classes, crules, thr)
        keep = [i for i in range(len(masks)) if i not in rem]
        masks = [masks[i] for i in keep]; scores = [scores[i] for i in keep]
        classes = [classes[i] for i in keep]; idx = [idx[i] for i in keep]
    return masks, scores, classes, idx
