"""Generate the committed golden vectors by running the UNMODIFIED reference (imported from /root/reference, see
refharness.py) on seeded synthetic inputs.  Run once in the build container:  python tests/golden/make_golden.py

Outputs (small, committed):
  measure_golden.npz : polygons -> reference findContours vertices + calculate_measurements values (src/utils/measurements.py:114)
  dedup_golden.npz   : head outputs -> reference deduplicate_masks_smart kept indices (src/functions/inference.py:2552),
                       filter_by_overlap_rules / filter_by_containment_rules removed sets (src/utils/spatial_constraints.py:192,:280),
                       greedy in-order de-dup with iou() (inference.py:422,:1453-1459), quirk KATs Q1/Q2
  paste_golden.npz   : head outputs -> bit-packed masks of the Detectron2 paste restatement (torch-CPU grid_sample; Detectron2 itself is
                       absent, so this file pins the restatement + installed torch, not the reference)
  misc_golden.npz    : rle_encoding (src/utils/mask_utils.py:17), is_edge_mask / generate_tiles_with_overlap (inference.py:2522,:2488),
                       postprocess_masks_universal (:1739), postprocess_masks (mask_utils.py:38)
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import refharness  # noqa: E402
from deepemia_b200 import synthetic as syn  # noqa: E402
from oracle import d2_paste  # noqa: E402

R = refharness.load_reference()
KEYS = ["major_axis_length", "minor_axis_length", "eccentricity", "Length", "Width", "CircularED", "Aspect_Ratio",
        "Circularity", "Chords", "Feret_diam", "Roundness", "Sphericity"]


def golden_measure():
    H = W = 512
    rng = np.random.default_rng(1000)
    polys = [np.round(p).astype(np.int32) for p in syn.particle_field(rng, 64, H, W)]
    verts, vstart, vals, inst, um_list, f32_flags = [], [0], [], [], [], None
    for i, p in enumerate(polys):
        m = np.zeros((H, W), np.uint8)
        cv2.fillPoly(m, [p], 255)
        um = [0.5, 1.0, 0.123][i % 3]
        for c in cv2.findContours(m, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)[0]:
            if cv2.contourArea(c) < 5:
                continue
            ref = R.measurements.calculate_measurements(c, m, um_pix=um, pixelsPerMetric=1)
            verts.append(c[:, 0, :].astype(np.int32))
            vstart.append(vstart[-1] + len(c))
            vals.append([float(ref[k]) for k in KEYS])
            f32_flags = [isinstance(ref[k], np.float32) for k in KEYS]
            inst.append(i)
            um_list.append(um)
    np.savez_compressed(os.path.join(HERE, "measure_golden.npz"), H=H, W=W,
                        poly_pts=np.concatenate(polys), poly_start=np.cumsum([0] + [len(p) for p in polys]),
                        verts=np.concatenate(verts), vstart=np.array(vstart), vals=np.array(vals, np.float64),
                        inst=np.array(inst), um=np.array(um_list), f32_flags=np.array(f32_flags), keys=np.array(KEYS))
    print("measure_golden:", len(vals), "contours")


def _ref_dedup_indices(ml, sl, cl, thr):
    rm, rs, rc = R.inference.deduplicate_masks_smart(ml, sl, cl, iou_threshold=thr)
    ids = {id(m): i for i, m in enumerate(ml)}
    return [ids[id(m)] for m in rm]


def golden_dedup():
    H = W = 256
    out = {}
    rules = syn.POLYHIPES_RULES
    for g in range(3):
        probs, boxes, scores, classes = syn.synthetic_heads(2000 + g, 48 + 8 * g, H, W, duplicate_frac=0.6, rmin=5, rmax=18, margin=20)
        masks = d2_paste.paste_masks_in_image(probs, boxes, (H, W))
        ml = [m for m in masks]; sl = [np.float32(s) for s in scores]; cl = [int(c) for c in classes]
        out[f"probs{g}"] = probs.astype(np.float16); out[f"boxes{g}"] = boxes; out[f"scores{g}"] = scores; out[f"classes{g}"] = classes
        for thr in (0.4, 0.7):
            out[f"smart{g}_{int(thr * 10)}"] = np.array(_ref_dedup_indices(ml, sl, cl, thr), np.int32)
        _, _, _, rem = R.spatial_constraints.filter_by_overlap_rules(ml, sl, cl, rules['overlap_rules'])
        out[f"overlap_removed{g}"] = np.array(sorted(rem), np.int32)
        _, _, _, rem = R.spatial_constraints.filter_by_containment_rules(ml, sl, cl, rules['containment_rules'], 0.95)
        out[f"contain_removed{g}"] = np.array(sorted(rem), np.int32)
        _, _, _, rem = R.spatial_constraints.filter_by_containment_rules(ml, sl, cl, {1: 0}, 0.5)
        out[f"contain50_removed{g}"] = np.array(sorted(rem), np.int32)
        # greedy in-order de-dup (inference.py:1453-1459) with the reference's iou()
        kept, kept_i = [], []
        for i, m in enumerate(ml):
            if not any(R.inference.iou(m, u) > 0.5 for u in kept):
                kept.append(m); kept_i.append(i)
        out[f"inorder{g}"] = np.array(kept_i, np.int32)
        out[f"pair_iou{g}"] = np.array([[R.inference.iou(ml[a], ml[b]), R.inference.calculate_iou(ml[a], ml[b]),
                                         R.spatial_constraints.calculate_iou(ml[a], ml[b]),
                                         R.spatial_constraints.calculate_containment(ml[a], ml[b])]
                                        for a in range(0, 12) for b in range(0, 12)], np.float64)

    def disc(x, y, r=8):
        m = np.zeros((128, 128), np.uint8); cv2.circle(m, (x, y), r, 1, -1); return m.astype(bool)
    kat = [([disc(20, 90), disc(20, 90)], [0.9, 0.8]), ([disc(90, 20), disc(90, 20)], [0.9, 0.8]),
           ([disc(64, 64)] * 3, [0.7, 0.8, 0.9]), ([disc(64, 64)] * 3, [0.9, 0.8, 0.7])]
    for k, (ms, sc) in enumerate(kat):
        ms = [m.copy() for m in ms]
        out[f"quirk{k}"] = np.array(_ref_dedup_indices(ms, [np.float32(s) for s in sc], [0] * len(ms), 0.4), np.int32)
    line = np.zeros((128, 128), bool); line[60, 10:100] = True
    out["thin_line_kept"] = np.array(_ref_dedup_indices([line], [np.float32(0.9)], [0], 0.4), np.int32)
    np.savez_compressed(os.path.join(HERE, "dedup_golden.npz"), H=H, W=W, **out)
    print("dedup_golden written")


def golden_paste():
    H, W = 200, 260
    probs, boxes, scores, classes = syn.synthetic_heads(3000, 24, H, W, rmin=4, rmax=25, margin=10)
    extra = np.array([[-10, -10, 30, 25], [W - 20, H - 15, W + 30, H + 9], [5, 5, 5, 40], [0, 0, W, H], [50.25, 60.5, 51.0, 61.25]], np.float32)
    eprobs = np.random.default_rng(5).random((len(extra), 28, 28)).astype(np.float16).astype(np.float32)
    probs = np.concatenate([probs, eprobs]); boxes = np.concatenate([boxes, extra])
    out = {"probs": probs.astype(np.float16), "boxes": boxes}
    for tag, (sx, sy) in {"a": (1.0, 1.0), "b": (1.28, 0.77)}.items():
        b, keep = d2_paste.detector_postprocess_boxes(boxes, sx, sy, H, W)
        masks = d2_paste.paste_masks_in_image(probs[keep], b[keep], (H, W))
        out[f"keep_{tag}"] = keep
        out[f"bits_{tag}"] = np.packbits(masks, axis=-1, bitorder="little")
    np.savez_compressed(os.path.join(HERE, "paste_golden.npz"), H=H, W=W, **out)
    print("paste_golden written")


def golden_misc():
    rng = np.random.default_rng(4000)
    H = W = 160
    polys = syn.particle_field(rng, 10, H, W, rmin=6, rmax=20, margin=25)
    masks = syn.masks_from_polys(polys, H, W)
    # holes + spurs so that fill-holes / opening have work to do
    for k, m in enumerate(masks):
        ys, xs = np.nonzero(m)
        cy, cx = int(ys.mean()), int(xs.mean())
        if k % 2 == 0:
            m[cy - 1:cy + 2, cx - 1:cx + 2] = 0
        if k % 3 == 0:
            m[ys.min() - 3:ys.min(), cx] = 1
    out = {"masks": np.packbits(np.stack(masks).astype(bool), axis=-1, bitorder="little"), "H": H, "W": W}
    rle = [np.array(R.mask_utils.rle_encoding(m), np.int64) for m in masks]
    out["rle"] = np.concatenate(rle); out["rle_start"] = np.cumsum([0] + [len(r) for r in rle])
    img = np.zeros((H, W, 3), np.uint8)
    for small in (True, False):
        res = []
        for m in masks:
            r = R.inference.postprocess_masks_universal(np.array([m.astype(bool)]), np.array([0.9]), img, 0, small)
            res.append(r[0] if r else np.zeros((H, W), bool))
        out[f"universal_{'small' if small else 'large'}"] = np.packbits(np.stack(res), axis=-1, bitorder="little")
        out[f"universal_{'small' if small else 'large'}_kept"] = np.array([len(R.inference.postprocess_masks_universal(
            np.array([m.astype(bool)]), np.array([0.9]), img, 0, small)) for m in masks])
    pm = R.mask_utils.postprocess_masks(np.stack(masks).astype(bool), np.linspace(0.9, 0.6, len(masks)).astype(np.float32), img, 5)
    out["postprocess_masks"] = np.packbits(np.stack(pm).astype(bool), axis=-1, bitorder="little")
    pp = R.inference.process_masks_parallel(pm)
    out["process_masks_parallel"] = np.packbits(np.stack(pp).astype(bool), axis=-1, bitorder="little")
    out["is_edge"] = np.array([R.inference.is_edge_mask(m.astype(bool), H, 0.25) for m in masks])
    big = np.zeros((700, 900, 3), np.uint8)
    tiles = R.inference.generate_tiles_with_overlap(big, 256, 0.125)
    out["tiles_xy"] = np.array([(x, y) for _, x, y in tiles], np.int32)
    out["tiles_8192"] = len(R.inference.generate_tiles_with_overlap(np.zeros((8192, 8192, 3), np.uint8), 1024, 0.125))
    np.savez_compressed(os.path.join(HERE, "misc_golden.npz"), **out)
    print("misc_golden written")


if __name__ == "__main__":
    golden_measure()
    golden_dedup()
    golden_paste()
    golden_misc()
