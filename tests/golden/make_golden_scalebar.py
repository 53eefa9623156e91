"""Golden vectors of row f3: the UNMODIFIED reference detect_scale_bar (src/utils/scalebar_ocr.py:72) on the synthetic frames of
scalebar_cases.py, with the EasyOCR reader replaced by one that returns the case's caption box (EasyOCR is absent from this
image and outside the path).  Run once in the build container:  python tests/golden/make_golden_scalebar.py
Output: scalebar_golden.npz (psum, um_pix per case + the thresholds / ROIs the reference read from its own config)."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import refharness  # noqa: E402
import scalebar_cases as sc  # noqa: E402

refharness.load_reference()
from src.utils import scalebar_ocr as ref  # noqa: E402
from src.utils.config import get_config  # noqa: E402


class FakeReader:
    result = []

    def __init__(self, *a, **k):
        pass

    def readtext(self, image, detail=1, paragraph=False):
        return FakeReader.result


ref.easyocr.Reader = FakeReader
cfg = get_config()
used = {"scale_bar_rois": cfg.get("scale_bar_rois", {}), "scalebar_thresholds": cfg.get("scalebar_thresholds", {})}
default_roi = ref.get_scalebar_roi_for_dataset(None)
out = {"config_json": np.array(json.dumps(used))}
for name in sc.CASES:
    image, ocr, kw = sc.build(name, default_roi)
    FakeReader.result = ocr
    psum, um_pix = ref.detect_scale_bar(image.copy(), **kw)
    out[name + "/psum"] = np.array(str(psum))
    out[name + "/um_pix"] = np.array(float(um_pix), np.float64)
    print(f"{name:28s} psum={psum!s:6s} um_pix={float(um_pix)!r}")
np.savez_compressed(os.path.join(HERE, "scalebar_golden.npz"), **out)
