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


class FakeHeadPredictor:
    """A deterministic stand-in for the Mask R-CNN heads (Detectron2 is absent): the head outputs of an image are a
    pure function of the image bytes (CRC32 -> seed), so the reference functions (fed the Detectron2-paste oracle of these heads)
    and the CUDA mirror (fed the raw heads) see the same detections for every image, tile, and rescaled image they derive.
    Test / golden-vector infrastructure."""

    def __init__(self, base_seed=0, n=28, duplicate_frac=0.35, rmin=5.0, rmax=15.0, margin=6, input_scale=1.0, zero_score=False):
        self.base_seed, self.n, self.dup, self.rmin, self.rmax, self.margin = base_seed, n, duplicate_frac, rmin, rmax, margin
        self.input_scale, self.zero_score = input_scale, zero_score
        self.calls = 0

    def raw_heads(self, image):
        """(probs [n,28,28] f32, boxes [n,4] f32 in model-input coordinates, scores, classes int64, (in_h, in_w))."""
        import zlib
        H, W = image.shape[:2]
        seed = (zlib.crc32(np.ascontiguousarray(image).tobytes()) ^ (self.base_seed * 2654435761)) & 0x7FFFFFFF
        n = self.n + int(seed % 7)
        r_hi = min(self.rmax, max(self.rmin + 1.0, min(H, W) / 6.0))
        probs, boxes, scores, classes = synthetic_heads(seed, n, H, W, duplicate_frac=self.dup, rmin=self.rmin, rmax=r_hi,
                                                        margin=self.margin)
        rng = np.random.default_rng(seed + 1)
        # a few degenerate / out-of-frame boxes exercise Boxes.nonempty() and the clipping
        boxes[0] = [W + 5.0, 3.0, W + 20.0, 18.0]
        boxes[1, 2] = boxes[1, 0]
        boxes[2] += np.array([-boxes[2, 0] - 4.0, 0.0, -boxes[2, 0] - 4.0, 0.0], np.float32)      # hangs over the left edge
        if self.zero_score:
            scores[3] = 0.0
        in_h, in_w = int(round(H * self.input_scale)), int(round(W * self.input_scale))
        boxes = boxes * np.array([in_w / W, in_h / H, in_w / W, in_h / H], np.float32)
        _ = rng
        self.calls += 1
        return probs, boxes.astype(np.float32), scores.astype(np.float32), classes.astype(np.int64), (in_h, in_w)

    def heads(self, image):
        """The pre-paste hook of deepemia_b200.functions.inference (CUDA tensors)."""
        import torch
        from .functions.inference import HeadOutputs
        probs, boxes, scores, classes, in_size = self.raw_heads(image)
        dev = torch.device("cuda", torch.cuda.current_device())
        t = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=dev)
        return HeadOutputs(t(probs), t(boxes), t(scores), t(classes), in_size)
