"""The scale-bar parity cases shared by make_golden_scalebar.py (reference side), tests/test_oracle_golden.py (oracle side) and
tests/test_gpu_scalebar.py (CUDA side): synthetic SEM frames with an info strip (bar + caption) and the OCR result a reader would
return for them (EasyOCR itself is outside the path: SURVEY.md §2 row 5)."""
import cv2
import numpy as np

BOTTOM_ROI = {"x_start_factor": 0.5, "y_start_factor": 0.86, "width_factor": 0.5, "height_factor": 0.12}
WIDE_ROI = {"x_start_factor": 0.05, "y_start_factor": 0.8, "width_factor": 0.9, "height_factor": 0.18}


def _roi_rect(h, w, roi):
    x0, y0 = int(w * roi["x_start_factor"]), int(h * roi["y_start_factor"])
    return x0, y0, min(int(x0 + w * roi["width_factor"]), w), min(int(y0 + h * roi["height_factor"]), h)


def make_frame(seed, h, w, roi, bars, caption="500 nm", caption_at=(0.28, 0.8), level=255, strip=20, noise=12, tilt=0, extra_lines=0):
    """BGR uint8 frame: textured micrograph, a dark info strip over the ROI, white bars (x0 fraction, y fraction, length px, thickness px)
    inside it and a caption.  Returns (image, ocr_result) with ocr_result in EasyOCR's readtext(detail=1) format, ROI coordinates."""
    rng = np.random.default_rng(seed)
    base = rng.integers(30, 200, (h // 6 + 1, w // 6 + 1), dtype=np.uint8)
    img = cv2.resize(base, (w, h), interpolation=cv2.INTER_CUBIC)
    img = np.clip(img.astype(np.int32) + rng.integers(-noise, noise + 1, (h, w)), 0, 255).astype(np.uint8)
    x0, y0, x1, y1 = _roi_rect(h, w, roi)
    rw, rh = x1 - x0, y1 - y0
    img[y0:y1, x0:x1] = np.clip(strip + rng.integers(-4, 5, (rh, rw)), 0, 255).astype(np.uint8)
    for fx, fy, length, thick in bars:
        bx, by = x0 + int(fx * rw), y0 + int(fy * rh)
        if tilt:
            cv2.line(img, (bx, by), (bx + length, by + tilt), int(level), thick)
        else:
            img[by:by + thick, bx:bx + length] = level
    for k in range(extra_lines):
        p = (x0 + int(rng.integers(0, rw)), y0 + int(rng.integers(0, rh)))
        q = (x0 + int(rng.integers(0, rw)), y0 + int(rng.integers(0, rh)))
        cv2.line(img, p, q, int(rng.integers(120, 256)), int(rng.integers(1, 3)))
    ocr = []
    if caption is not None:
        org = (x0 + int(caption_at[0] * rw), y0 + int(caption_at[1] * rh))
        (tw, th), bl = cv2.getTextSize(caption, cv2.FONT_HERSHEY_SIMPLEX, 0.5, 1)
        cv2.putText(img, caption, org, cv2.FONT_HERSHEY_SIMPLEX, 0.5, 255, 1)
        bx0, by0 = org[0] - x0, org[1] - th - y0
        ocr.append(([[bx0, by0], [bx0 + tw, by0], [bx0 + tw, by0 + th + bl], [bx0, by0 + th + bl]], caption, 0.93))
    return cv2.cvtColor(img, cv2.COLOR_GRAY2BGR), ocr


# name -> (frame kwargs, detect_scale_bar kwargs).  roi None = the configured default ROI (a 5 % strip near the top right).
CASES = {
    "plain_bar": (dict(seed=1, h=768, w=1024, roi=BOTTOM_ROI, bars=[(0.15, 0.3, 200, 4)]), dict(roi_config=BOTTOM_ROI)),
    "thin_bar": (dict(seed=2, h=768, w=1024, roi=BOTTOM_ROI, bars=[(0.2, 0.35, 150, 2)], caption="200 nm"), dict(roi_config=BOTTOM_ROI)),
    "split_bar_merged": (dict(seed=3, h=768, w=1024, roi=BOTTOM_ROI, bars=[(0.1, 0.3, 90, 3), (0.1 + 102 / 512, 0.3, 110, 3)], caption="1 um"),
                         dict(roi_config=BOTTOM_ROI)),
    "split_bar_far_apart": (dict(seed=4, h=768, w=1024, roi=BOTTOM_ROI, bars=[(0.1, 0.3, 80, 3), (0.1 + 130 / 512, 0.3, 120, 3)]),
                            dict(roi_config=BOTTOM_ROI)),
    "two_bars_two_rows": (dict(seed=5, h=768, w=1024, roi=BOTTOM_ROI, bars=[(0.15, 0.2, 120, 3), (0.2, 0.5, 220, 3)]), dict(roi_config=BOTTOM_ROI)),
    "bar_far_from_text": (dict(seed=6, h=900, w=1600, roi=WIDE_ROI, bars=[(0.7, 0.3, 200, 4)], caption_at=(0.05, 0.8)), dict(roi_config=WIDE_ROI)),
    "dim_bar": (dict(seed=7, h=768, w=1024, roi=BOTTOM_ROI, bars=[(0.15, 0.3, 200, 4)], level=90), dict(roi_config=BOTTOM_ROI)),
    "dim_bar_low_threshold": (dict(seed=7, h=768, w=1024, roi=BOTTOM_ROI, bars=[(0.15, 0.3, 200, 4)], level=90),
                              dict(roi_config=BOTTOM_ROI, intensity_threshold=40)),
    "no_caption": (dict(seed=8, h=768, w=1024, roi=BOTTOM_ROI, bars=[(0.15, 0.3, 200, 4)], caption=None), dict(roi_config=BOTTOM_ROI)),
    "caption_without_digits": (dict(seed=9, h=768, w=1024, roi=BOTTOM_ROI, bars=[(0.15, 0.3, 200, 4)], caption="SEM HV"), dict(roi_config=BOTTOM_ROI)),
    "bar_at_roi_edge": (dict(seed=10, h=768, w=1024, roi=BOTTOM_ROI, bars=[(0.01, 0.3, 200, 4)]), dict(roi_config=BOTTOM_ROI)),
    "tilted_bar": (dict(seed=11, h=768, w=1024, roi=BOTTOM_ROI, bars=[(0.15, 0.3, 200, 3)], tilt=12), dict(roi_config=BOTTOM_ROI)),
    "steep_line": (dict(seed=12, h=768, w=1024, roi=BOTTOM_ROI, bars=[(0.3, 0.1, 60, 3)], tilt=55), dict(roi_config=BOTTOM_ROI)),
    "clutter": (dict(seed=13, h=768, w=1024, roi=BOTTOM_ROI, bars=[(0.15, 0.3, 180, 4)], extra_lines=6), dict(roi_config=BOTTOM_ROI)),
    "clutter_wide": (dict(seed=14, h=900, w=1600, roi=WIDE_ROI, bars=[(0.1, 0.45, 300, 5)], caption="10 um", caption_at=(0.12, 0.85), extra_lines=12),
                     dict(roi_config=WIDE_ROI, proximity_threshold=180)),
    "short_bar": (dict(seed=15, h=768, w=1024, roi=BOTTOM_ROI, bars=[(0.2, 0.3, 26, 3)]), dict(roi_config=BOTTOM_ROI)),
    "noisy_strip": (dict(seed=16, h=768, w=1024, roi=BOTTOM_ROI, bars=[(0.15, 0.3, 240, 4)], strip=70, noise=40), dict(roi_config=BOTTOM_ROI)),
    "default_roi_top_right": (dict(seed=17, h=1200, w=1600, roi=None, bars=[(0.2, 0.25, 180, 4)], caption="2 um", caption_at=(0.25, 0.9)), dict()),
    "default_roi_dataset": (dict(seed=18, h=1200, w=1600, roi=None, bars=[(0.3, 0.2, 120, 3)], caption="100", caption_at=(0.3, 0.92)),
                            dict(dataset_name="polyhipes")),
    "explicit_proximity": (dict(seed=19, h=768, w=1024, roi=BOTTOM_ROI, bars=[(0.15, 0.3, 200, 4)]),
                           dict(roi_config=BOTTOM_ROI, proximity_threshold=30)),
    "explicit_intensity": (dict(seed=2, h=768, w=1024, roi=BOTTOM_ROI, bars=[(0.2, 0.35, 150, 2)], caption="200 nm"),
                           dict(roi_config=BOTTOM_ROI, intensity_threshold=90)),
    "merge_gap_exact": (dict(seed=22, h=768, w=1024, roi=BOTTOM_ROI, bars=[(0.1, 0.3, 100, 3), (0.1 + 115 / 512, 0.3, 100, 3)], caption="1 um"),
                        dict(roi_config=BOTTOM_ROI)),
    "rows_within_y_tolerance": (dict(seed=23, h=768, w=1024, roi=BOTTOM_ROI, bars=[(0.1, 0.3, 100, 3), (0.1 + 108 / 512, 0.34, 100, 3)]),
                                dict(roi_config=BOTTOM_ROI)),
    "small_frame": (dict(seed=20, h=300, w=400, roi=WIDE_ROI, bars=[(0.1, 0.3, 120, 3)], caption="50 nm", caption_at=(0.2, 0.9)), dict(roi_config=WIDE_ROI)),
    "roi_overhanging": (dict(seed=21, h=768, w=1024, roi={"x_start_factor": 0.6, "y_start_factor": 0.9, "width_factor": 0.8, "height_factor": 0.3},
                             bars=[(0.2, 0.3, 150, 4)], caption_at=(0.2, 0.8)),
                        dict(roi_config={"x_start_factor": 0.6, "y_start_factor": 0.9, "width_factor": 0.8, "height_factor": 0.3})),
}


def build(name, default_roi):
    fk, dk = CASES[name]
    fk = dict(fk)
    if fk["roi"] is None:
        fk["roi"] = default_roi
    image, ocr = make_frame(**fk)
    return image, ocr, dict(dk)
