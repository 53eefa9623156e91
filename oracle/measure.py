"""Oracle: measurement loop + ``calculate_measurements`` (numpy/OpenCV restatement; test infrastructure only).

Follows the reference:
  * measurement loop          src/functions/inference.py:1148-1253 (binarise, findContours EXTERNAL/SIMPLE, per-contour
                              ``contourArea >= max(5, H*W*0.000005*0.05)`` gate, one row per surviving contour — Q12)
  * ``calculate_measurements`` src/utils/measurements.py:114-233 (Q3, Q4, Q13, Q14)
  * ``imutils.perspective.order_points`` / ``imutils.grab_contours`` (imutils is unpinned and absent: SURVEY B.3)
  * CSV header                src/functions/inference.py:987-1010
The per-instance JPEG dump (:1153-1162) and ``gc.collect()`` (:1253) have no effect on results and are omitted.
"""
import cv2
import numpy as np
from scipy.spatial import distance as dist

CSV_HEADER = [
    "Instance_ID", "Class", "Class_Name", "Major axis length", "Minor axis length", "Eccentricity", "C. Length",
    "C. Width", "Circular eq. diameter", "Aspect ratio", "Circularity", "Chord length", "Ferret diameter",
    "Roundness", "Sphericity", "Contrast d10", "Contrast d50", "Contrast d90", "Detected scale bar", "File name",
]
MEASUREMENT_KEYS = ["major_axis_length", "minor_axis_length", "eccentricity", "Length", "Width", "CircularED",
                    "Aspect_Ratio", "Circularity", "Chords", "Feret_diam", "Roundness", "Sphericity"]


def order_points(pts):
    xs = pts[np.argsort(pts[:, 0]), :]
    left, right = xs[:2, :], xs[2:, :]
    left = left[np.argsort(left[:, 1]), :]
    tl, bl = left
    d = dist.cdist(tl[np.newaxis], right, "euclidean")[0]
    br, tr = right[np.argsort(d)[::-1], :]
    return np.array([tl, tr, br, bl], dtype="float32")


def _mid(a, b):
    return ((a[0] + b[0]) * 0.5, (a[1] + b[1]) * 0.5)


def calculate_measurements(c, um_pix=1.0, pixelsPerMetric=1.0, gray=None, mask=None):
    area = cv2.contourArea(c)
    perimeter = cv2.arcLength(c, True)
    box = np.array(cv2.boxPoints(cv2.minAreaRect(c)), dtype="int")
    tl, tr, br, bl = order_points(box)
    dA = dist.euclidean(_mid(tl, tr), _mid(bl, br))
    dB = dist.euclidean(_mid(tl, bl), _mid(tr, br))
    dimA, dimB = dA / pixelsPerMetric, dB / pixelsPerMetric
    dimArea, dimPer = area / pixelsPerMetric, perimeter / pixelsPerMetric
    aspect = max(dimB, dimA) / min(dimA, dimB) if (dimA and dimB) != 0 else 0
    out = {
        "Length": min(dimA, dimB) * um_pix,
        "Width": max(dimA, dimB) * um_pix,
        "CircularED": np.sqrt(4 * area / np.pi) * um_pix,
        "Aspect_Ratio": aspect,
        "Chords": cv2.arcLength(c, True) * um_pix,
        "Roundness": 1 / aspect if aspect != 0 else 0,
        "Sphericity": (2 * np.sqrt(np.pi * dimArea)) / dimPer * um_pix if dimPer != 0 else 0,
        "Circularity": 4 * np.pi * (dimArea / (dimPer) ** 2) * um_pix if dimPer != 0 else 0,
        "Feret_diam": max(dimA, dimB) * um_pix,
    }
    if len(c) >= 5:
        (_, _), (major_axis, minor_axis), _ = cv2.fitEllipse(c)
        a, b = (major_axis / 2.0, minor_axis / 2.0) if major_axis > minor_axis else (minor_axis / 2.0, major_axis / 2.0)
        out["eccentricity"] = np.sqrt(1 - (b ** 2 / a ** 2)) if a != 0 else 0
        out["major_axis_length"] = major_axis / pixelsPerMetric * um_pix
        out["minor_axis_length"] = minor_axis / pixelsPerMetric * um_pix
    else:
        out["eccentricity"] = out["major_axis_length"] = out["minor_axis_length"] = 0
    d10 = d50 = d90 = None
    if gray is not None and mask is not None:
        px = gray[mask > 0]
        if len(px) > 0:
            hist, edges = np.histogram(px, bins=256, range=(0, 255), density=True)
            cdf = np.cumsum(hist)
            cdf /= cdf[-1]
            d10, d50, d90 = (np.interp(q, cdf, edges[:-1]) for q in (0.10, 0.50, 0.90))
    out["contrast_d10"], out["contrast_d50"], out["contrast_d90"] = d10, d50, d90
    return out


def min_contour_area(image_shape):
    # inference.py:1178-1184
    return max(5, image_shape[0] * image_shape[1] * 0.000005 * 0.05)


def measure_masks(masks, classes, image_shape, um_pix, test_img="img", class_names=None, psum="0", image=None,
                  measure_contrast_distribution=False):
    """Rows of measurements_results.csv for one image (inference.py:1148-1253)."""
    rows = []
    min_area = min_contour_area(image_shape)
    gray = None
    if measure_contrast_distribution and image is not None:
        gray = cv2.cvtColor(image, cv2.COLOR_BGR2GRAY) if image.ndim == 3 else image.copy()
    for instance_id, (mask, cls) in enumerate(zip(masks, classes), 1):
        binary = (np.asarray(mask) > 0).astype(np.uint8) * 255
        cnts = cv2.findContours(binary, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
        cnts = cnts[0] if len(cnts) == 2 else cnts[1]
        for c in cnts:
            if cv2.contourArea(c) < min_area:
                continue
            m = calculate_measurements(c, um_pix=um_pix, pixelsPerMetric=1, gray=gray, mask=binary)
            name = class_names[cls] if class_names is not None and cls < len(class_names) else f"class_{cls}"
            rows.append([f"{test_img}_{instance_id}", cls, name] + [m[k] for k in MEASUREMENT_KEYS] +
                        [m["contrast_d10"], m["contrast_d50"], m["contrast_d90"], psum, test_img])
    return rows
