"""Synthetic micrograph instances: star-polygon particles as full-frame masks or as Mask R-CNN head
outputs (28x28 probability map + box).  Host-side numpy/OpenCV; used by tests, bench.py and the golden generator."""
import cv2
import numpy as np

MASK_SIDE = 28


def star_polygon(rng, cx, cy, r):
    k = int(rng.integers(5, 13))
    ang = np.sort(rng.uniform(0, 2 * np.pi, k))
    rad = r * rng.uniform(0.6, 1.0, k)
    return np.stack([cx + rad * np.cos(ang), cy + rad * np.sin(ang)], 1)


def particle_field(rng, n, H, W, rmin=8.0, rmax=30.0, margin=40):
    """n star polygons (float vertices) with centres uniform inside the margin."""
    polys = []
    for _ in range(n):
        r = rng.uniform(rmin, rmax)
        cx = rng.uniform(margin, W - margin)
        cy = rng.uniform(margin, H - margin)
        polys.append(star_polygon(rng, cx, cy, r))
    return polys


def masks_from_polys(polys, H, W):
    out = []
    for p in polys:
        m = np.zeros((H, W), np.uint8)
        cv2.fillPoly(m, [np.round(p).astype(np.int32)], 1)
        out.append(m)
    return out


def head_from_poly(rng, poly):
    """(prob 28x28 float32 fp16-representable, box xyxy float32) for one particle."""
    x0, y0 = poly.min(0) - rng.uniform(0, 2, 2)
    x1, y1 = poly.max(0) + rng.uniform(0, 2, 2)
    sx, sy = MASK_SIDE / (x1 - x0), MASK_SIDE / (y1 - y0)
    q = np.stack([(poly[:, 0] - x0) * sx, (poly[:, 1] - y0) * sy], 1)
    m = np.zeros((MASK_SIDE, MASK_SIDE), np.float32)
    cv2.fillPoly(m, [np.round(q * 16).astype(np.int32)], 1.0, lineType=cv2.LINE_AA, shift=4)
    m = cv2.GaussianBlur(m, (0, 0), 1.0)
    m = np.clip(m, 0, 1).astype(np.float16).astype(np.float32)
    return m, np.array([x0, y0, x1, y1], np.float32)


def distinct_scores(rng, n, lo=0.05, hi=1.0):
    s = rng.uniform(lo, hi, n).astype(np.float32)
    while len(np.unique(s)) != n:
        s = rng.uniform(lo, hi, n).astype(np.float32)
    return s


def synthetic_heads(seed, n, H, W, duplicate_frac=0.0, **kw):
    """n head outputs for one H x W tile: probs (n,28,28) f32, boxes (n,4) f32, scores (n,) f32 distinct, classes (n,) i32.
    duplicate_frac > 0 re-detects that fraction of particles with a small box jitter (work for the de-dup step)."""
    rng = np.random.default_rng(seed)
    n_base = max(1, int(round(n / (1.0 + duplicate_frac))))
    polys = particle_field(rng, n_base, H, W, **kw)
    while len(polys) < n:
        src = polys[int(rng.integers(0, n_base))]
        polys.append(src + rng.uniform(-1.5, 1.5, 2))
    probs = np.zeros((n, MASK_SIDE, MASK_SIDE), np.float32)
    boxes = np.zeros((n, 4), np.float32)
    for i, p in enumerate(polys):
        probs[i], boxes[i] = head_from_poly(rng, p)
    scores = distinct_scores(rng, n)
    classes = (rng.random(n) < 0.5).astype(np.int32)
    return probs, boxes, scores, classes


# spatial rules of config/datasets/polyhipes_tommy.yaml:41-58 (class 1 must lie inside class 0; max IoU 0.30 / 0.50)
POLYHIPES_RULES = {
    'enabled': True,
    'containment_rules': {1: 0},
    'containment_threshold': 0.95,
    'overlap_rules': {0: {'allow_overlap': False, 'allow_touch': True, 'max_iou_threshold': 0.30},
                      1: {'allow_overlap': False, 'allow_touch': True, 'max_iou_threshold': 0.50}},
}
