"""CPU: the host build of the device algorithms (deepemia_b200/csrc/core/*.cuh compiled by g++ into tests/hostsim) against
OpenCV / torch and the reference-generated golden vectors.  These are the exact functions the sm_100a kernels execute."""
import ctypes
import os
import subprocess

import cv2
import numpy as np
import pytest

from oracle import d2_paste, measure

HERE = os.path.dirname(os.path.abspath(__file__))
HS_DIR = os.path.join(HERE, "hostsim")
HS = os.path.join(HS_DIR, "libemia_hostsim.so")


@pytest.fixture(scope="module")
def L():
    srcs = [os.path.join(HS_DIR, "hostsim.cpp")] + [os.path.join(HERE, "..", "deepemia_b200", "csrc", "core", f)
                                                    for f in os.listdir(os.path.join(HERE, "..", "deepemia_b200", "csrc", "core"))]
    if not os.path.exists(HS) or any(os.path.getmtime(s) > os.path.getmtime(HS) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++", "hostsim.cpp", "-o", HS], cwd=HS_DIR)
    lib = ctypes.CDLL(HS)
    lib.sim_contour_area.restype = ctypes.c_double
    lib.sim_arc_length.restype = ctypes.c_double
    lib.sim_last_perimeter.restype = ctypes.c_double
    return lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _pack(c):
    c = c[:, 0, :].astype(np.uint32)
    return (c[:, 0] | (c[:, 1] << 16)).astype(np.uint32)


def _contours(L, mask):
    mask = np.ascontiguousarray(mask.astype(np.uint8))
    H, W = mask.shape
    pts = np.zeros(H * W + 16, np.uint32); cs = np.zeros(H * W // 2 + 2, np.int32); npts = ctypes.c_int()
    n = L.sim_find_contours(_p(mask), H, W, _p(pts), len(pts), _p(cs), len(cs) - 1, ctypes.byref(npts))
    return [pts[cs[k]:cs[k + 1]].copy() for k in range(n - 1, -1, -1)]


def _shapes(rng, n, H=200, W=260):
    for _ in range(n):
        m = np.zeros((H, W), np.uint8)
        u = rng.random()
        if u < 0.3:
            cv2.ellipse(m, (int(rng.integers(40, W - 40)), int(rng.integers(40, H - 40))), (int(rng.integers(1, 38)), int(rng.integers(1, 38))),
                        float(rng.uniform(0, 180)), 0, 360, 1, -1)
        elif u < 0.4:
            x0, y0 = rng.integers(5, W - 60), rng.integers(5, H - 60)
            m[y0:y0 + rng.integers(1, 50), x0:x0 + rng.integers(1, 50)] = 1
        else:
            k = rng.integers(3, 13); a = np.sort(rng.uniform(0, 2 * np.pi, k)); r = rng.uniform(0.5, 1, k) * rng.uniform(2, 38)
            cx, cy = rng.integers(40, W - 40), rng.integers(40, H - 40)
            cv2.fillPoly(m, [np.stack([cx + r * np.cos(a), cy + r * np.sin(a)], 1).astype(np.int32)], 1)
        yield m


def test_contours_area_perimeter_vs_opencv(L):
    rng = np.random.default_rng(1)
    imgs = [(rng.random((rng.integers(1, 40), rng.integers(1, 70))) < d).astype(np.uint8) for d in (0.1, 0.3, 0.5, 0.7, 0.9) for _ in range(120)]
    imgs += list(_shapes(rng, 300))
    for m in imgs:
        ref = cv2.findContours(np.ascontiguousarray(m), cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)[0]
        got = _contours(L, m)
        assert len(ref) == len(got)
        if len(ref):     # running arcLength accumulated while following = cv2.arcLength of contours[0] (the last discovered)
            assert L.sim_last_perimeter() == cv2.arcLength(ref[0], True)
        for r, p in zip(ref, got):
            assert np.array_equal(_pack(r), p)
            assert L.sim_contour_area(_p(p), len(p)) == cv2.contourArea(r)
            assert L.sim_arc_length(_p(p), len(p)) == cv2.arcLength(r, True)


def test_min_area_rect_and_box_points_bit_exact(L):
    rng = np.random.default_rng(3)
    n = 0
    for m in _shapes(rng, 1500):
        for c in cv2.findContours(m, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)[0]:
            p = _pack(c); rect = np.zeros(11, np.float32); box = np.zeros(8, np.float32)
            L.sim_min_area_rect(_p(p), len(p), 0, _p(rect), _p(box))
            rr = cv2.minAreaRect(c)
            assert np.array_equal(np.array([rr[0][0], rr[0][1], rr[1][0], rr[1][1], rr[2]], np.float32), rect[:5])
            assert np.array_equal(cv2.boxPoints(rr).reshape(-1), box)
            n += 1
    assert n >= 1500


def test_fit_ellipse_axes(L):
    rng = np.random.default_rng(4)
    tot = exact = 0
    for m in _shapes(rng, 1500):
        for c in cv2.findContours(m, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)[0]:
            if len(c) < 5 or cv2.contourArea(c) < 5:
                continue
            p = _pack(c); out = np.zeros(5, np.float32)
            L.sim_fit_ellipse(_p(p), len(p), _p(out))
            (_, _), (w, h), _ = cv2.fitEllipse(c)
            ref = np.array([w, h], np.float32)
            tot += 1
            exact += np.array_equal(ref, out[2:4])
            np.testing.assert_allclose(out[2:4], ref, rtol=1e-5)
    assert exact >= 0.99 * tot


def test_fit_ellipse_small_and_degenerate_shapes(L):
    """Tiny / thin blobs: 5-vertex contours (OpenCV's direct solver), two-pixel-wide strips (degenerate conic: the centre system
    is singular -> minimum-norm solve), shapes with |dx| == |dy| for every vertex (rank-deficient stage 2).  cv2.fitEllipse
    itself is non-reproducible on inputs where it perturbs the points with an RNG (call it twice: two answers) — those are
    skipped; of the rest at most 1 % may differ (axes that are pure rounding noise, e.g. 3.7e7 px for a two-row strip)."""
    rng = np.random.default_rng(0)
    tot = bad = nondet = n5 = 0
    for _ in range(6000):
        m = np.zeros((16, 24), np.uint8)
        h, w, y, x = rng.integers(1, 5), rng.integers(2, 12), rng.integers(2, 8), rng.integers(2, 8)
        m[y:y + h, x:x + w] = 1
        for _ in range(rng.integers(0, 5)):
            m[rng.integers(y, y + h), rng.integers(x, x + w)] = rng.integers(0, 2)
        for c in cv2.findContours(m * 255, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)[0]:
            if len(c) < 5 or cv2.contourArea(c) < 5:
                continue
            ref, ref2 = cv2.fitEllipse(c), cv2.fitEllipse(c)
            if ref != ref2:
                nondet += 1
                continue
            if not np.all(np.isfinite(ref[1])):
                continue
            p = _pack(c); out = np.zeros(5, np.float32)
            L.sim_fit_ellipse(_p(p), len(p), _p(out))
            tot += 1
            n5 += len(c) == 5
            bad += not np.allclose(out[2:4], np.array(ref[1], np.float32), rtol=1e-5, atol=1e-9)
    assert tot > 1000 and n5 > 50
    assert bad <= 0.01 * tot, (bad, tot, nondet)


def test_measure_contour_vs_reference_golden(L):
    g = np.load(os.path.join(HERE, "golden", "measure_golden.npz"))
    for k in range(len(g["vals"])):
        v = g["verts"][g["vstart"][k]:g["vstart"][k + 1]].astype(np.uint32)
        p = (v[:, 0] | (v[:, 1] << 16)).astype(np.uint32)
        rec = np.zeros(16)
        L.sim_measure_contour(_p(p), len(p), ctypes.c_double(float(g["um"][k])), _p(rec))
        np.testing.assert_allclose(rec[:12], g["vals"][k], rtol=1e-5, atol=0)
        # integer-exact quantities
        c = v.astype(np.int32).reshape(-1, 1, 2)
        assert rec[12] == cv2.contourArea(c) and rec[13] == cv2.arcLength(c, True) and rec[14] == len(v)


def test_paste_core_bit_exact_vs_torch(L):
    rng = np.random.default_rng(0)
    H, W = 129, 161
    for t in range(60):
        prob = [rng.random((28, 28)).astype(np.float32), rng.random((28, 28)).astype(np.float16).astype(np.float32),
                (rng.integers(0, 257, (28, 28)) / 256).astype(np.float32)][t % 3]
        x0, y0 = rng.uniform(-20, W - 5), rng.uniform(-20, H - 5)
        box = np.array([x0, y0, x0 + rng.uniform(0.5, 90), y0 + rng.uniform(0.5, 90)], np.float32)
        sx, sy = (1.0, 1.0) if t % 2 else (float(rng.uniform(0.5, 2)), float(rng.uniform(0.5, 2)))
        ref, _, _, _ = d2_paste.predictor_instances(prob[None], box[None], np.array([0.9], np.float32), np.array([0]), sx, sy, H, W)
        out = np.zeros((H, W), np.uint8); reg = np.zeros(4, np.int32)
        v = L.sim_paste(_p(prob), _p(box), ctypes.c_float(sx), ctypes.c_float(sy), H, W, _p(out), _p(reg))
        assert v == len(ref)
        if v:
            assert np.array_equal(ref[0], out.astype(bool))


def test_moments_bit_exact_vs_opencv(L):
    """cv2.moments(mask) (src/functions/inference.py:1101): the 10 raw moments and the 3 second-order central moments bit for
    bit; third-order central and normalised moments (differences of numbers ~1e6 x larger) to 1e-9 of their natural scale —
    on random shapes, also far from the origin of a large frame."""
    from deepemia_b200.engine import MOMENT_NAMES
    rng = np.random.default_rng(11)
    n = 0
    for m in _shapes(rng, 150):
        for big in (False, True):
            mask = m
            if big:
                mask = np.zeros((1400, 2100), np.uint8)
                oy, ox = int(rng.integers(0, 1200)), int(rng.integers(0, 1840))
                mask[oy:oy + m.shape[0], ox:ox + m.shape[1]] = m
            mask = np.ascontiguousarray(mask)
            out = np.zeros(24, np.float64)
            L.sim_moments(_p(mask), mask.shape[0], mask.shape[1], _p(out))
            ref = cv2.moments(mask)
            want = np.array([ref[k] for k in MOMENT_NAMES])
            assert np.array_equal(out[:13], want[:13]), (n, [(k, a, b) for k, a, b in zip(MOMENT_NAMES, out, want) if a != b][:4])
            ext = max(mask.shape)                                             # coordinates up to `ext`
            assert np.allclose(out[13:17], want[13:17], rtol=1e-9, atol=1e-12 * want[0] * ext ** 3)
            assert np.allclose(out[17:20], want[17:20], rtol=1e-12, atol=0) and np.allclose(out[20:], want[20:], rtol=1e-6, atol=1e-8 + 1e-13 * ext ** 3 / max(want[0], 1.0) ** 1.5)
            n += 1
    empty = np.zeros((20, 40), np.uint8)
    out = np.ones(24, np.float64)
    L.sim_moments(_p(empty), 20, 40, _p(out))
    assert np.array_equal(out, np.array([cv2.moments(empty)[k] for k in MOMENT_NAMES]))


def test_wavelength_mirror_matches_reference_text():
    """rgb_to_hsv / hue_to_wavelength / rgb_to_wavelength (src/utils/measurements.py:32-111): known answers worked out from the
    reference's formulas (hue halved OpenCV-style, 620 - 170/270 * hue)."""
    from deepemia_b200.utils import measurements as M
    assert M.rgb_to_hsv(255, 0, 0) == (0.0, 255.0, 255.0)
    assert M.rgb_to_hsv(0, 255, 0) == (60.0, 255.0, 255.0)
    assert M.rgb_to_hsv(0, 0, 255) == (120.0, 255.0, 255.0)
    assert M.rgb_to_hsv(0, 0, 0) == (0.0, 0.0, 0.0) and M.rgb_to_hsv(7, 7, 7)[:2] == (0.0, 0.0)
    assert M.rgb_to_hsv(255, 0, 255)[0] == 150.0            # magenta: -60 + 360 = 300 -> halved
    assert M.rgb_to_wavelength(255, 0, 0) == 620
    assert M.rgb_to_wavelength(0, 255, 0) == 620 - 170 / 270 * 60.0
    assert M.rgb_to_wavelength(0, 0, 255) == 620 - 170 / 270 * 120.0
    assert abs(M.rgb_to_wavelength(30, 200, 120) - (620 - 170 / 270 * (60 * ((120 / 255 - 30 / 255) / (200 / 255 - 30 / 255)) + 120) / 2)) < 1e-9


# ---- row f3: scale-bar line detection core (core/emia_scalebar.cuh) vs OpenCV -------------------------------------------------
def _sb_frame(rng, H, W):
    img = cv2.GaussianBlur(rng.integers(0, 90, (H, W, 3)).astype(np.uint8), (5, 5), 0)
    for _ in range(int(rng.integers(1, 5))):
        x0, y0 = int(rng.integers(5, W - 60)), int(rng.integers(5, H - 10))
        cv2.rectangle(img, (x0, y0), (x0 + int(rng.integers(20, min(200, W - x0 - 2))), y0 + int(rng.integers(1, 6))), (255, 255, 255), -1)
    if rng.random() < 0.5:
        cv2.putText(img, "500 nm", (int(rng.integers(0, W // 2)), int(rng.integers(15, H))), cv2.FONT_HERSHEY_SIMPLEX, 0.6, (255, 255, 255), 2)
    if rng.random() < 0.5:
        cv2.line(img, (int(rng.integers(0, W)), int(rng.integers(0, H))), (int(rng.integers(0, W)), int(rng.integers(0, H))), (220, 220, 220),
                 int(rng.integers(1, 4)))
    return img


def test_scalebar_core_bit_exact_vs_opencv(L):
    """BGR2GRAY, Canny, HoughLinesP (OpenCV's random visiting order included) and the thickness-2 line mask of the device core are
    bit-identical to cv2 on synthetic strips, pure-noise images (dense edges: hundreds of lines) and non-default parameters."""
    from deepemia_b200.engine import hough_tables
    rng = np.random.default_rng(0)
    n_lines = 0
    for it in range(150):
        H, W = int(rng.integers(20, 120)), int(rng.integers(80, 400))
        img = _sb_frame(rng, H, W) if it % 3 else rng.integers(0, 256, (H, W, 3)).astype(np.uint8)
        gray = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
        g2 = np.zeros_like(gray)
        L.sim_bgr2gray(_p(np.ascontiguousarray(img)), H * W, _p(g2))
        assert (g2 == gray).all()
        lo, hi = (50, 150) if it % 2 else (int(rng.integers(10, 100)), int(rng.integers(100, 300)))
        e = cv2.Canny(gray, lo, hi, apertureSize=3)
        e2 = np.zeros_like(e)
        L.sim_canny(_p(gray), H, W, lo, hi, _p(e2))
        assert (e == e2).all(), it
        if it % 2:
            rho, theta, thr, ml, mg = 1.0, np.pi / 180, 50, 20, 10
        else:
            rho, theta = float(rng.choice([1.0, 2.0, 0.5])), float(rng.choice([np.pi / 180, np.pi / 90, np.pi / 360]))
            thr, ml, mg = int(rng.integers(10, 60)), int(rng.integers(5, 40)), int(rng.integers(0, 15))
        lines = cv2.HoughLinesP(e, rho, theta, threshold=thr, minLineLength=ml, maxLineGap=mg)
        ref = np.zeros((0, 4), np.int32) if lines is None else lines[:, 0, :]
        trig, numangle, numrho = hough_tables(W, H, rho, theta)
        out = np.zeros((8192, 4), np.int32)
        nl = L.sim_hough_lines_p(_p(e), H, W, _p(trig), numangle, numrho, thr, ml, mg, _p(out), len(out))
        assert nl == len(ref) and (out[:nl] == ref).all(), (it, nl, len(ref))
        n_lines += nl
        for x1, y1, x2, y2 in list(ref[:8]) + [rng.integers(0, [W, H, W, H]) for _ in range(4)]:
            m = np.zeros((H, W), np.uint8)
            cv2.line(m, (int(x1), int(y1)), (int(x2), int(y2)), 255, 2)
            m2 = np.zeros((H, W), np.uint8)
            L.sim_thick_line(_p(m2), H, W, int(x1), int(y1), int(x2), int(y2))
            assert (m == m2).all(), (it, x1, y1, x2, y2)
    assert n_lines > 2000
