"""TEST INFRASTRUCTURE (oracle): CPU restatement of the reference's scale-bar line detection, src/utils/scalebar_ocr.py:72-463,
with OpenCV itself doing the image work (cv2.cvtColor, cv2.Canny, cv2.HoughLinesP, cv2.line + cv2.mean — the reference's own
third-party calls).  The OCR result is an input (EasyOCR's readtext(detail=1) format).  Pinned against the unmodified reference by
tests/golden/scalebar_golden.npz (tests/golden/make_golden_scalebar.py).  Only tests/ may import this module."""
import re
from math import sqrt

import cv2
import numpy as np


def roi_rect(h, w, roi_config):
    """scalebar_ocr.py:125-128 (the slice at :139 clamps the far edges)."""
    x0 = int(w * roi_config["x_start_factor"])
    y0 = int(h * roi_config["y_start_factor"])
    x1 = int(x0 + w * roi_config["width_factor"])
    y1 = int(y0 + h * roi_config["height_factor"])
    return x0, y0, min(x1, w), min(y1, h)


def text_box(ocr_result):
    """scalebar_ocr.py:159-194: first detection with a digit -> (psum, centre) in ROI coordinates."""
    for bbox, text, _ in ocr_result or []:
        clean = re.sub("[^0-9]", "", text)
        if clean:
            xs, ys = [p[0] for p in bbox], [p[1] for p in bbox]
            return clean, ((int(min(xs)) + int(max(xs))) // 2, (int(min(ys)) + int(max(ys))) // 2)
    return "0", None


def merge_group(group):
    """scalebar_ocr.py:430-463."""
    if len(group) == 1:
        return group[0]
    all_x = [s["x1"] for s in group] + [s["x2"] for s in group]
    all_y = [s["y1"] for s in group] + [s["y2"] for s in group]
    x1, x2 = min(all_x), max(all_x)
    y1 = y2 = int(sum(all_y) / len(all_y))
    tl = sum(s["length"] for s in group)
    return {"x1": x1, "y1": y1, "x2": x2, "y2": y2, "length": sqrt((x2 - x1) ** 2 + (y2 - y1) ** 2),
            "intensity": sum(s["intensity"] * s["length"] for s in group) / tl,
            "dist_to_text": sum(s["dist_to_text"] * s["length"] for s in group) / tl}


def merge_collinear(segments, max_gap=15, y_tolerance=5):
    """scalebar_ocr.py:376-427."""
    if not segments:
        return []
    ss = sorted(segments, key=lambda s: min(s["x1"], s["x2"]))
    merged, cur = [], [ss[0]]
    for seg in ss[1:]:
        last = cur[-1]
        gap = min(seg["x1"], seg["x2"]) - max(last["x1"], last["x2"])
        if gap <= max_gap and abs((seg["y1"] + seg["y2"]) / 2 - (last["y1"] + last["y2"]) / 2) <= y_tolerance:
            cur.append(seg)
        else:
            merged.append(merge_group(cur))
            cur = [seg]
    merged.append(merge_group(cur))
    return merged


def detect_scale_bar(image, ocr_result, roi_config, intensity_threshold=200, proximity_threshold=50, thresholds=None, details=None):
    """scalebar_ocr.py:72-373 with the config look-ups resolved by the caller: thresholds = the `scalebar_thresholds` mapping."""
    th = thresholds or {}
    if "intensity" in th and intensity_threshold == 200:
        intensity_threshold = th["intensity"]
    if "proximity" in th and proximity_threshold == 50:
        proximity_threshold = th["proximity"]
    merge_gap, min_len, margin = th.get("merge_gap", 15), th.get("min_line_length", 30), th.get("edge_margin_factor", 0.1)
    h, w = image.shape[:2]
    x0, y0, x1, y1 = roi_rect(h, w, roi_config)
    gray = cv2.cvtColor(image[y0:y1, x0:x1].copy(), cv2.COLOR_BGR2GRAY)
    rh, rw = gray.shape[:2]
    xm, ym = int(rw * margin), int(rh * margin)
    psum, centre = text_box(ocr_result)
    edges = cv2.Canny(gray, 50, 150, apertureSize=3)
    longest, max_length = None, 0
    lines = None
    if centre:
        lines = cv2.HoughLinesP(edges, 1, np.pi / 180, threshold=50, minLineLength=20, maxLineGap=10)
        raw = []
        for pts in (lines if lines is not None else []):
            ax, ay, bx, by = pts[0]
            ang = abs(np.arctan2(by - ay, bx - ax) * 180 / np.pi)
            if ang > 10 and ang < 170:
                continue
            if min(ax, bx) < xm or max(ax, bx) > rw - xm or min(ay, by) < ym or max(ay, by) > rh - ym:
                continue
            m = np.zeros_like(gray, dtype=np.uint8)
            cv2.line(m, (ax, ay), (bx, by), 255, 2)
            c = ((ax + bx) // 2, (ay + by) // 2)
            raw.append({"x1": ax, "y1": ay, "x2": bx, "y2": by, "length": sqrt((bx - ax) ** 2 + (by - ay) ** 2),
                        "intensity": cv2.mean(gray, mask=m)[0], "dist_to_text": sqrt((c[0] - centre[0]) ** 2 + (c[1] - centre[1]) ** 2)})
        merged = merge_collinear(raw, merge_gap)
        for s in merged:
            near = (min(s["x1"], s["x2"]) < xm or max(s["x1"], s["x2"]) > rw - xm or min(s["y1"], s["y2"]) < ym or max(s["y1"], s["y2"]) > rh - ym)
            if s["dist_to_text"] < proximity_threshold and s["intensity"] > intensity_threshold and s["length"] > min_len and not near:
                if s["length"] > max_length:
                    max_length, longest = s["length"], (s["x1"], s["y1"], s["x2"], s["y2"])
        if details is not None:
            details.update(raw=raw, merged=merged)
    if details is not None:
        details.update(gray=gray, edges=edges, lines=lines, longest=longest, scale_len=max_length, roi=(x0, y0, x1, y1))
    if longest:
        return psum, (float(psum) / max_length if max_length > 0 else 1.0)
    return "0", 1
