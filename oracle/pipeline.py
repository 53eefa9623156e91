"""Oracle: the per-tile hot path assembled from the reference's own steps (test infrastructure only).

Order follows ``run_inference`` for one image (src/functions/inference.py): predictor output (Detectron2 paste,
oracle.d2_paste) -> ``deduplicate_masks_smart`` (:859, iou 0.7) -> ``apply_spatial_constraints`` (:868) -> measurement loop
(:1148-1253).  Returns the kept original head indices (in final list order) and the CSV rows.
"""
import numpy as np

from . import d2_paste, dedup, measure, spatial


def run_tile(probs, boxes, scores, classes, H, W, um_pix=0.5, rules=None, dedup_iou=0.7, scale_x=1.0, scale_y=1.0,
             test_img="tile", psum="500", class_names=None):
    b, keep = d2_paste.detector_postprocess_boxes(boxes, scale_x, scale_y, H, W)
    orig = np.nonzero(keep)[0]
    masks = d2_paste.paste_masks_in_image(np.asarray(probs, np.float32)[keep], b[keep], (H, W)) if len(orig) else []
    ml = [m for m in masks]
    sl = [np.float32(s) for s in np.asarray(scores)[keep]]
    cl = [int(c) for c in np.asarray(classes)[keep]]
    m2, s2, c2, idx = dedup.deduplicate_masks_smart(ml, sl, cl, iou_threshold=dedup_iou, return_indices=True)
    m3, s3, c3, idx2 = spatial.apply_spatial_constraints(m2, s2, c2, rules)
    final = [int(orig[idx[i]]) for i in idx2]
    rows = measure.measure_masks(m3, c3, (H, W), um_pix, test_img=test_img, class_names=class_names, psum=psum)
    return final, rows, m3
