"""Oracle: the inference FLOWS assembled from the reference's own steps (numpy / OpenCV / scipy on full-frame masks, exactly
the representation and the per-mask Python loops of the reference).  TEST INFRASTRUCTURE ONLY (tests/, bench.py's CPU legs).

Follows src/functions/inference.py of the reference:
  * ``run_class_specific_inference``   :1353-1461  (single predictor)
  * ``tile_based_inference_pipeline``  :2299-2485
  * per-image body of ``run_inference`` :776-905 (per class tile pipeline -> deduplicate_masks_smart(0.7) :859 ->
    apply_spatial_constraints :868) and the measurement loop :1148-1253
  * BASELINE config 4: run_ensemble_inference :1464-1598 merged with the multi-scale pass
    (process_single_scale :1987-2066 without its iterative re-detection, score-sorted iou() de-dup :1955-1978)
Pinned against the reference golden (tests/golden/flows_golden.npz) by tests/test_oracle_golden.py.

An ``instances`` argument is what ``predictor(image)["instances"]`` holds after Detectron2's paste:
(masks [N, H, W] bool, scores [N] float32, classes [N] int64)."""
import cv2
import numpy as np

from . import d2_paste, dedup, measure, morphology, spatial, tiles


def heads_to_instances(probs, boxes, scores, classes, H, W, scale_x=1.0, scale_y=1.0):
    """detector_postprocess + paste_masks_in_image (oracle.d2_paste) of raw head outputs."""
    masks, s, c, _ = d2_paste.predictor_instances(probs, boxes, scores, classes, scale_x, scale_y, H, W)
    return np.asarray(masks), s.astype(np.float32), c.astype(np.int64)


def run_class_specific_inference(instances, target_class, small_classes, confidence_threshold=0.3, iou_threshold=0.7,
                                 min_size=None, parallel=True):
    pred_masks, pred_scores, pred_classes = instances
    class_mask = pred_classes == target_class
    cm, cs = pred_masks[class_mask], pred_scores[class_mask]
    conf = cs >= confidence_threshold
    fm, fs = cm[conf], cs[conf]
    if len(fm) == 0:
        return [], [], []
    is_small = target_class in small_classes
    if min_size is None:
        min_size = 5 if is_small else 25
    processed = morphology.postprocess_masks(fm, fs, fm.shape[1:], min_crys_size=min_size)
    if len(processed) > 2 and parallel:
        processed = morphology.process_masks_parallel(processed)
    if not processed:
        return [], [], []
    thr = 0.5 if is_small else iou_threshold
    m, s, c, _ = dedup.greedy_inorder_dedup(processed, fs, target_class, thr)
    return m, s, c


def tile_based_inference_pipeline(full_instances, tile_instances, tile_xy, image_hw, target_class, small_classes,
                                  confidence_threshold, tile_size=512, overlap_ratio=0.1, iou_threshold=0.7,
                                  edge_filter_enabled=True, min_size=None, parallel=True):
    """tile_instances[t]: instances of the UPSCALED tile t; tile_xy[t] = (x_offset, y_offset)."""
    h, w = image_hw
    fm, fs, fc = run_class_specific_inference(full_instances, target_class, small_classes, confidence_threshold, iou_threshold,
                                              min_size, parallel)
    am, asc, ac = list(fm), list(fs), list(fc)
    for inst, (x, y) in zip(tile_instances, tile_xy):
        tm, ts_, tc = run_class_specific_inference(inst, target_class, small_classes, confidence_threshold, iou_threshold, min_size,
                                                   parallel)
        for mask, score, cls in zip(tm, ts_, tc):
            g = tiles.back_project(mask, tile_size, tile_size, int(x), int(y), h, w, tile_size, overlap_ratio, edge_filter_enabled)
            if g is None:
                continue
            am.append(g); asc.append(score); ac.append(cls)
    return dedup.deduplicate_masks_smart(am, asc, ac, iou_threshold=0.4)


def infer_image(full_instances, tile_instances, tile_xy, image_hw, params, small_classes, tile_size, overlap_ratio, rules=None,
                edge_filter_enabled=True, cross_class_iou=0.7):
    """params: list of dicts(target_class, confidence_threshold, iou_threshold, min_size).  -> (masks, scores, classes)."""
    am, asc, ac = [], [], []
    for p in params:
        m, s, c = tile_based_inference_pipeline(full_instances, tile_instances, tile_xy, image_hw, p["target_class"], small_classes,
                                                p["confidence_threshold"], tile_size, overlap_ratio, p.get("iou_threshold", 0.7),
                                                edge_filter_enabled, p.get("min_size"))
        am += list(m); asc += list(s); ac += list(c)
    if not am:
        return [], [], []
    m, s, c = dedup.deduplicate_masks_smart(am, asc, ac, iou_threshold=cross_class_iou)
    if rules and rules.get("enabled", False) and len(m):
        m, s, c, _ = spatial.apply_spatial_constraints(m, s, c, rules)
    return m, s, c


def ensemble_multiscale(instances, weights, image_hw, params, small_classes, rules=None, sorted_iou=0.4, cross_class_iou=0.7):
    """instances[scale][model] = instances of model `model` on the image rescaled by `scale`."""
    h, w = image_hw
    area0 = h * w
    am, asc, ac = [], [], []
    for p in params:
        t = p["target_class"]
        is_small = t in small_classes
        base_min = max(3, int(area0 * 0.000005)) if is_small else max(25, int(area0 * 0.0001))
        cm, cs = [], []
        for scale, per_model in instances.items():
            for (masks, scores, classes), weight in zip(per_model, weights):
                sel = (classes == t) & (scores >= p["confidence_threshold"])
                for mask, score in zip(masks[sel], scores[sel]):
                    cleaned = morphology.postprocess_masks_universal(np.array([mask]), mask.shape, is_small,
                                                                     min_crys_size=int(base_min * (scale ** 2)))
                    if not cleaned:
                        continue
                    m = cleaned[0]
                    if scale != 1.0:
                        m = cv2.resize(m.astype(np.uint8), (w, h), interpolation=cv2.INTER_NEAREST).astype(bool)
                    cm.append(m); cs.append(score * weight)
        um, us = [], []
        if cs:
            for idx in np.argsort(cs)[::-1]:
                if not any(dedup.iou(cm[idx], e) > sorted_iou for e in um):
                    um.append(cm[idx]); us.append(cs[idx])
        m, s, c = dedup.deduplicate_masks_smart(um, us, [t] * len(um), iou_threshold=p.get("iou_threshold", 0.7))
        am += list(m); asc += list(s); ac += list(c)
    if not am:
        return [], [], []
    m, s, c = dedup.deduplicate_masks_smart(am, asc, ac, iou_threshold=cross_class_iou)
    if rules and rules.get("enabled", False) and len(m):
        m, s, c, _ = spatial.apply_spatial_constraints(m, s, c, rules)
    return m, s, c


def measure_rows(masks, classes, image_hw, um_pix):
    """The float columns of the measurement rows (src/functions/inference.py:1148-1253) of the final masks."""
    rows = measure.measure_masks(masks, classes, image_hw, um_pix, test_img="x", class_names=None, psum="0")
    return [[float(v) for v in r[3:15]] for r in rows]
