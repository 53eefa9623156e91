"""CPU: host-side pieces of the drop-in layer that need no GPU (pure Python mirrors of reference code)."""

import numpy as np

from deepemia_b200.functions import inference as inf


def test_class_color_legend_format(tmp_path):
    """class_color_legend.txt of run_inference (src/functions/inference.py:1302-1314): header, one line per class, BGR colour
    table of :972-981 printed as an RGB tuple, colours repeat after eight classes."""
    names = [f"c{i}" for i in range(10)]
    path = inf.write_class_color_legend(str(tmp_path), names)
    lines = open(path).read().splitlines()
    assert lines[0] == "Class Color Legend:" and lines[1] == "=================="
    assert lines[2] == "Class 0 (c0): RGB(0, 255, 0)"          # green, BGR (0, 255, 0)
    assert lines[3] == "Class 1 (c1): RGB(0, 0, 255)"          # blue, BGR (255, 0, 0)
    assert lines[9] == "Class 7 (c7): RGB(0, 165, 255)"        # BGR (255, 165, 0)
    assert lines[10] == "Class 8 (c8): RGB(0, 255, 0)"         # wraps around
    assert len(lines) == 12


def test_csv_header_is_the_reference_schema():
    """The 20 columns of measurements_results.csv (src/functions/inference.py:987-1010), in order."""
    assert inf.CSV_HEADER == [
        "Instance_ID", "Class", "Class_Name", "Major axis length", "Minor axis length", "Eccentricity", "C. Length", "C. Width",
        "Circular eq. diameter", "Aspect ratio", "Circularity", "Chord length", "Ferret diameter", "Roundness", "Sphericity",
        "Contrast d10", "Contrast d50", "Contrast d90", "Detected scale bar", "File name"]


def test_generate_tiles_with_overlap_matches_reference_arithmetic():
    """stride = int(ts * (1 - ov)); tiles at range(0, h, stride) x range(0, w, stride), zero-padded to ts x ts
    (src/functions/inference.py:2488-2519; SURVEY Appendix C: 8192^2, 1024, 0.125 -> 100 tiles)."""
    img = np.zeros((300, 420, 3), np.uint8)
    img[..., 0] = (np.arange(420) % 251)[None, :]
    tiles = inf.generate_tiles_with_overlap(img, 128, 0.25)
    stride = int(128 * 0.75)
    exp = [(x, y) for y in range(0, 300, stride) for x in range(0, 420, stride)]
    assert [(x, y) for _, x, y in tiles] == exp
    for t, x, y in tiles:
        assert t.shape == (128, 128, 3)
        h, w = min(128, 300 - y), min(128, 420 - x)
        assert np.array_equal(t[:h, :w], img[y:y + h, x:x + w]) and not t[h:].any() and not t[:, w:].any()
    big = np.zeros((8192, 8192, 1), np.uint8)
    assert len(inf.generate_tiles_with_overlap(big, 1024, 0.125)) == 100
