"""Scale-bar detection, interface mirror of the reference's src/utils/scalebar_ocr.py (row f3).

`detect_scale_bar` keeps the reference's signature, return convention `(psum: str, um_pix: float)` and selection rules
(scalebar_ocr.py:72-373); `detect_scale_bars` does the same for a batch of equally sized micrographs with ONE launch per stage.
What runs where:

* libemia.so (engine.scalebar_edges / hough_lines_p / line_means): BGR2GRAY of the ROI, cv2.Canny(50, 150), cv2.HoughLinesP(1,
  pi/180, 50, 20, 10) and the mean grey level under every line's thickness-2 mask — bit-identical to the OpenCV calls at
  scalebar_ocr.py:140, :200, :207-214, :247-249;
* host, over the handful of lines that come back: the horizontal / margin / proximity / intensity filters, the merge of collinear
  segments (:376-463) and the choice of the longest line — the reference's own scalar logic;
* the OCR itself (EasyOCR, a CRNN) is a registered reader `ocr(gray_roi) -> [(bbox, text, confidence), ...]` (EasyOCR's
  `readtext(detail=1)` format).  Without one, `easyocr` is imported if present (ONE Reader for the process: the reference builds a
  new Reader per call, :150); if it is absent the result is the reference's "EasyOCR failed" branch (no text -> ("0", 1)).
"""
import logging
import re
from math import sqrt
from typing import List

import numpy as np

from .. import engine

log = logging.getLogger("deepemia_b200.scalebar")

DEFAULT_ROI = {"x_start_factor": 0.7, "y_start_factor": 0.05, "width_factor": 1, "height_factor": 0.05}
_ROI_KEYS = ("x_start_factor", "y_start_factor", "width_factor", "height_factor")
_state = {"config": None, "ocr": None, "easyocr_reader": None}


class ScaleBarDetectionError(Exception):
    pass


def set_config_provider(fn):
    """fn(dataset_name) -> dict: the reference's get_config(dataset_name=...) (`scale_bar_rois`, `scalebar_thresholds`)."""
    _state["config"] = fn


def set_ocr_reader(fn):
    """fn(gray_roi: HxW uint8) -> [(bbox 4x2, text, confidence), ...] (EasyOCR readtext(detail=1, paragraph=False) format)."""
    _state["ocr"] = fn


def easyocr_available():
    import importlib.util
    return importlib.util.find_spec("easyocr") is not None


def _config(dataset_name):
    return (_state["config"](dataset_name) or {}) if _state["config"] else {}


def get_scalebar_roi_for_dataset(dataset_name: str = None) -> dict:
    """scalebar_ocr.py:29-69: dataset entry of `scale_bar_rois`, else its `default`, else the built-in default."""
    try:
        rois = _config(dataset_name).get("scale_bar_rois", {})
        if dataset_name and dataset_name in rois:
            return rois[dataset_name]
        return rois.get("default", DEFAULT_ROI)
    except Exception as e:                                   # noqa: BLE001 - the reference falls back on any config error
        log.error("Error loading scale bar ROI config: %s", e)
        return DEFAULT_ROI


def _read_text(gray_roi, ocr):
    ocr = ocr or _state["ocr"]
    if ocr is not None:
        return ocr(gray_roi)
    try:
        if _state["easyocr_reader"] is None:
            import easyocr
            _state["easyocr_reader"] = easyocr.Reader(["en"], verbose=False)
        return _state["easyocr_reader"].readtext(gray_roi, detail=1, paragraph=False)
    except Exception as e:                                   # noqa: BLE001 - scalebar_ocr.py:152-154
        log.error("EasyOCR failed: %s", e)
        return []


def _first_number(result):
    """(psum, text_box_center) of the first detection containing a digit (scalebar_ocr.py:159-194)."""
    for bbox, text, _ in result or []:
        digits = re.sub("[^0-9]", "", text)
        if digits:
            xs = [bbox[k][0] for k in range(4)]
            ys = [bbox[k][1] for k in range(4)]
            x_min, y_min, x_max, y_max = int(min(xs)), int(min(ys)), int(max(xs)), int(max(ys))
            return digits, ((x_min + x_max) // 2, (y_min + y_max) // 2), (x_min, y_min, x_max, y_max, text)
    return "0", None, None


def merge_segment_group(group: List[dict]) -> dict:
    """scalebar_ocr.py:430-463: leftmost / rightmost x, mean y (truncated), length-weighted intensity and distance."""
    if len(group) == 1:
        return group[0]
    xs = [s["x1"] for s in group] + [s["x2"] for s in group]
    ys = [s["y1"] for s in group] + [s["y2"] for s in group]
    x1, x2 = min(xs), max(xs)
    y = int(sum(ys) / len(ys))
    total = sum(s["length"] for s in group)
    return {"x1": x1, "y1": y, "x2": x2, "y2": y, "length": sqrt((x2 - x1) ** 2 + (y - y) ** 2),
            "intensity": sum(s["intensity"] * s["length"] for s in group) / total,
            "dist_to_text": sum(s["dist_to_text"] * s["length"] for s in group) / total, "line_idx": -1}


def merge_collinear_segments(segments: List[dict], max_gap: int = 15, angle_tolerance: int = 5, y_tolerance: int = 5) -> List[dict]:
    """scalebar_ocr.py:376-427: segments sorted by their left end; a segment joins the running group when its left end is within
    max_gap of the previous segment's right end and their mean heights differ by at most y_tolerance."""
    if not segments:
        return []
    order = sorted(segments, key=lambda s: min(s["x1"], s["x2"]))
    merged, group = [], [order[0]]
    for seg in order[1:]:
        last = group[-1]
        gap = min(seg["x1"], seg["x2"]) - max(last["x1"], last["x2"])
        y_offset = abs((seg["y1"] + seg["y2"]) / 2 - (last["y1"] + last["y2"]) / 2)
        if gap <= max_gap and y_offset <= y_tolerance:
            group.append(seg)
        else:
            merged.append(merge_segment_group(group))
            group = [seg]
    merged.append(merge_segment_group(group))
    return merged


def _thresholds(dataset_name, intensity_threshold, proximity_threshold):
    merge_gap, min_line_length, edge_margin_factor = 15, 30, 0.1
    try:
        th = _config(dataset_name).get("scalebar_thresholds", {})
        if "intensity" in th and intensity_threshold == 200:
            intensity_threshold = th["intensity"]
        if "proximity" in th and proximity_threshold == 50:
            proximity_threshold = th["proximity"]
        merge_gap = th.get("merge_gap", 15)
        min_line_length = th.get("min_line_length", 30)
        edge_margin_factor = th.get("edge_margin_factor", 0.1)
    except Exception as e:                                   # noqa: BLE001 - scalebar_ocr.py:113-114
        log.warning("Could not load thresholds from config: %s", e)
    return intensity_threshold, proximity_threshold, merge_gap, min_line_length, edge_margin_factor


def _roi_rect(h, w, roi_config):
    for key in _ROI_KEYS:
        if key not in roi_config:
            raise ScaleBarDetectionError(f"ROI config missing key: {key}")
    x_start = int(w * roi_config["x_start_factor"])
    y_start = int(h * roi_config["y_start_factor"])
    x_end = int(x_start + w * roi_config["width_factor"])
    y_end = int(y_start + h * roi_config["height_factor"])
    return x_start, y_start, x_end, y_end


def select_scale_line(lines, sums, text_box_center, roi_w, roi_h, intensity_threshold, proximity_threshold, merge_gap,
                      min_line_length, edge_margin_factor):
    """The reference's scalar logic over HoughLinesP's output (scalebar_ocr.py:216-303): lines [n,4] ints, sums [n,2] = (grey sum,
    pixel count) under each line.  Returns (longest_line or None, max_length, merged segment infos)."""
    x_margin, y_margin = int(roi_w * edge_margin_factor), int(roi_h * edge_margin_factor)

    def near_edge(x1, y1, x2, y2):
        return (min(x1, x2) < x_margin or max(x1, x2) > roi_w - x_margin or min(y1, y2) < y_margin or max(y1, y2) > roi_h - y_margin)

    raw = []
    for k, (x1, y1, x2, y2) in enumerate((int(a), int(b), int(c), int(d)) for a, b, c, d in lines):
        angle = abs(np.arctan2(y2 - y1, x2 - x1) * 180 / np.pi)
        if 10 < angle < 170 or near_edge(x1, y1, x2, y2):
            continue
        cx, cy = (x1 + x2) // 2, (y1 + y2) // 2
        total, count = int(sums[k][0]), int(sums[k][1])
        raw.append({"x1": x1, "y1": y1, "x2": x2, "y2": y2, "length": sqrt((x2 - x1) ** 2 + (y2 - y1) ** 2),
                    "intensity": total * (1.0 / count) if count else 0.0,          # cv2.mean(gray, mask)[0]
                    "dist_to_text": sqrt((cx - text_box_center[0]) ** 2 + (cy - text_box_center[1]) ** 2), "line_idx": k})
    merged = merge_collinear_segments(raw, merge_gap)
    longest, max_length = None, 0
    for seg in merged:
        seg["near_edge"] = near_edge(seg["x1"], seg["y1"], seg["x2"], seg["y2"])
        if (seg["dist_to_text"] < proximity_threshold and seg["intensity"] > intensity_threshold and seg["length"] > min_line_length
                and not seg["near_edge"] and seg["length"] > max_length):
            max_length = seg["length"]
            longest = (seg["x1"], seg["y1"], seg["x2"], seg["y2"])
    return longest, max_length, merged


def _draw_debug(image, rect, text_info, merged, longest, max_length):
    import cv2                                              # debugging overlay only (scalebar_ocr.py draw_debug)
    x0, y0, x1, y1 = rect
    cv2.rectangle(image, (x0, y0), (x1, y1), (0, 255, 0), 2)
    cv2.putText(image, "ROI", (x0, y0 - 10), cv2.FONT_HERSHEY_SIMPLEX, 0.5, (0, 255, 0), 2)
    if text_info is not None:
        tx0, ty0, tx1, ty1, text = text_info
        cv2.rectangle(image, (x0 + tx0, y0 + ty0), (x0 + tx1, y0 + ty1), (255, 0, 0), 2)
        cv2.putText(image, f"Text: {text}", (x0 + tx0, y0 + ty0 - 5), cv2.FONT_HERSHEY_SIMPLEX, 0.4, (255, 0, 0), 1)
    for k, seg in enumerate(merged):
        color = (128, 128, 128) if seg["near_edge"] else (255, 255, 0)
        cv2.line(image, (x0 + seg["x1"], y0 + seg["y1"]), (x0 + seg["x2"], y0 + seg["y2"]), color, 1)
        cv2.putText(image, f"M{k}: {seg['length']:.0f}px, I:{seg['intensity']:.0f}, D:{seg['dist_to_text']:.0f}"
                    + (" [EDGE]" if seg["near_edge"] else ""), (x0 + seg["x1"], y0 + seg["y1"] - 5), cv2.FONT_HERSHEY_SIMPLEX, 0.3, color, 1)
    if longest:
        cv2.line(image, (longest[0] + x0, longest[1] + y0), (longest[2] + x0, longest[3] + y0), (0, 0, 255), 3)
        cv2.putText(image, f"SELECTED: {max_length:.0f}px", (longest[0] + x0, longest[1] + y0 - 10), cv2.FONT_HERSHEY_SIMPLEX, 0.5,
                    (0, 0, 255), 2)
    else:
        cv2.putText(image, "SCALE BAR DETECTION FAILED", (x0, y0 + 30), cv2.FONT_HERSHEY_SIMPLEX, 0.7, (0, 0, 255), 2)


def detect_scale_bars(images, roi_config=None, intensity_threshold=200, proximity_threshold=50, dataset_name=None, draw_debug=False,
                      ocr=None, return_details=False):
    """detect_scale_bar for a batch: images = list of HxWx3 BGR uint8 arrays (one launch per stage and distinct frame size), or one
    [B,H,W,3] array / device tensor.
    Returns [(psum, um_pix)] (with return_details: also a dict per image with edges, lines, merged segments)."""
    if roi_config is None:
        roi_config = get_scalebar_roi_for_dataset(dataset_name)
    intensity_threshold, proximity_threshold, merge_gap, min_line_length, edge_margin_factor = _thresholds(
        dataset_name, intensity_threshold, proximity_threshold)
    import torch
    if isinstance(images, (list, tuple)):
        for im in images:
            if not isinstance(im, np.ndarray):
                raise ScaleBarDetectionError("Input image is not a numpy array.")
        if not images:
            return []
        shapes = {}
        for k, im in enumerate(images):
            shapes.setdefault(im.shape, []).append(k)
        if len(shapes) > 1:
            # frames of different sizes (a dataset mixing microscopes): one batch per size, results back in input order
            if return_details:
                raise ScaleBarDetectionError("return_details needs equally shaped images")
            out = [None] * len(images)
            for idx in shapes.values():
                res = detect_scale_bars([images[k] for k in idx], roi_config, intensity_threshold, proximity_threshold, dataset_name,
                                        draw_debug, ocr)
                for k, r in zip(idx, res):
                    out[k] = r
            return out
        B, (h, w) = len(images), images[0].shape[:2]
    elif isinstance(images, (np.ndarray, torch.Tensor)):
        B, h, w = (int(v) for v in images.shape[:3])
        if B == 0:
            return []
    else:
        raise ScaleBarDetectionError("Input image is not a numpy array.")
    x_start, y_start, x_end, y_end = _roi_rect(h, w, roi_config)
    x1c, y1c = min(x_end, w), min(y_end, h)                    # numpy slicing clamps the far edge (scalebar_ocr.py:139)
    if x1c <= x_start or y1c <= y_start:
        raise ScaleBarDetectionError(f"empty scale bar ROI x={x_start}:{x_end}, y={y_start}:{y_end}")
    # only the rows of the ROI travel to the device
    rows = np.stack([im[y_start:y1c] for im in images]) if isinstance(images, (list, tuple)) else images[:, y_start:y1c]
    gray, edges = engine.scalebar_edges(rows, (x_start, 0, x1c, y1c - y_start), 50, 150)
    roi_h, roi_w = int(gray.shape[1]), int(gray.shape[2])
    gray_h = gray.cpu().numpy()                                # the OCR reader's input (a few KB per image)
    texts = [_first_number(_read_text(gray_h[b], ocr)) for b in range(B)]
    out, details = [], []
    lines_h = n_h = sums_h = None
    if any(t[1] is not None for t in texts):
        cap = 1024
        while True:
            lines, n_lines = engine.hough_lines_p(edges, 1, np.pi / 180, threshold=50, min_line_length=20, max_line_gap=10, max_lines=cap)
            n_h = n_lines.cpu().numpy()
            if int(n_h.max()) <= cap:
                break
            cap = int(n_h.max())
        sums_h = engine.line_means(gray, lines, n_lines).cpu().numpy()
        lines_h = lines.cpu().numpy()
    for b in range(B):
        psum, center, text_info = texts[b]
        longest, max_length, merged = None, 0, []
        if center is not None:
            n = int(n_h[b])
            longest, max_length, merged = select_scale_line(lines_h[b, :n], sums_h[b, :n], center, roi_w, roi_h, intensity_threshold,
                                                            proximity_threshold, merge_gap, min_line_length, edge_margin_factor)
        if longest:
            um_pix = float(psum) / max_length if max_length > 0 else 1.0
        else:
            um_pix, psum = 1, "0"
            log.warning("No scale bar line detected near OCR text.")
        if draw_debug and isinstance(images, (list, tuple)):
            _draw_debug(images[b], (x_start, y_start, x_end, y_end), text_info, merged, longest, max_length)
        out.append((psum, um_pix))
        if return_details:
            details.append({"roi": (x_start, y_start, x1c, y1c), "lines": None if lines_h is None else lines_h[b, :int(n_h[b])].copy(),
                            "merged": merged, "longest": longest, "scale_len": max_length})
    if return_details:
        return out, {"gray": gray, "edges": edges, "per_image": details}
    return out


def detect_scale_bar(image, roi_config=None, intensity_threshold=200, proximity_threshold=50, dataset_name=None, draw_debug=False,
                     ocr=None):
    """scalebar_ocr.py:72-373, same signature (+ the optional OCR reader) and return value (psum: str, um_pix: float)."""
    if not isinstance(image, np.ndarray):
        raise ScaleBarDetectionError("Input image is not a numpy array.")
    return detect_scale_bars([image], roi_config, intensity_threshold, proximity_threshold, dataset_name, draw_debug, ocr)[0]
