"""GPU parity of the small reference-interface mirrors (deepemia_b200.utils.*, helper functions of functions/inference.py,
the measurement loop and the CSV writer) against the oracle restatements (which tests/test_oracle_golden.py pins to the reference)."""
import csv
import os

import cv2
import numpy as np
import pytest

from deepemia_b200 import synthetic as syn
from deepemia_b200.functions import inference as inf
from deepemia_b200.utils import mask_utils, measurements, spatial_constraints as sc, contrast
from oracle import dedup, measure as omeasure, morphology, spatial as ospatial, tiles

pytestmark = pytest.mark.gpu


def _masks(seed, n, H, W, dup=10):
    rng = np.random.default_rng(seed)
    polys = syn.particle_field(rng, n, H, W, rmin=6, rmax=22, margin=10)
    polys += [p + rng.uniform(-2.5, 2.5, 2) for p in polys[:dup]]
    return [m.astype(bool) for m in syn.masks_from_polys(polys, H, W)], rng


def test_helpers(cuda_device):
    ms, _ = _masks(1, 8, 120, 150, dup=3)
    e = np.zeros((120, 150), bool)
    assert inf.iou(ms[0], ms[8]) == dedup.iou(ms[0], ms[8]) and inf.iou(e, e) == 0
    assert inf.get_mask_bbox(ms[2]) == tuple(int(v) for v in dedup.get_mask_bbox(ms[2])) and inf.get_mask_bbox(e) is None
    b0, b8 = dedup.get_mask_bbox(ms[0]), dedup.get_mask_bbox(ms[8])
    assert inf.calculate_iou(ms[0], ms[8], b0, b8) == dedup.calculate_iou(ms[0], ms[8], b0, b8)
    assert sc.calculate_containment(ms[0], ms[8]) == ospatial.calculate_containment(ms[0], ms[8], b0, b8)
    for m in ms[:4] + [e]:
        assert inf.is_edge_mask(m, 120, 0.3) == tiles.is_edge_mask(m, 120, 0.3)
    img = np.random.default_rng(0).integers(0, 255, (130, 170, 3), dtype=np.uint8)
    got, ref = inf.generate_tiles_with_overlap(img, 64, 0.25), tiles.generate_tiles_with_overlap(img, 64, 0.25)
    assert len(got) == len(ref) and all(a[1:] == b[1:] and np.array_equal(a[0], b[0]) for a, b in zip(got, ref))
    assert mask_utils.rle_encoding(ms[1].astype(np.uint8)) == [int(v) for v in morphology.rle_encoding(ms[1].astype(np.uint8))]


def test_mask_cleanup_mirrors(cuda_device):
    H, W = 160, 200
    ms, rng = _masks(2, 14, H, W, dup=8)
    ring = np.zeros((H, W), np.uint8); cv2.circle(ring, (90, 80), 25, 1, 3); ms.insert(2, ring.astype(bool))
    image = np.zeros((H, W, 3), np.uint8)
    scores = syn.distinct_scores(rng, len(ms))
    ref = morphology.postprocess_masks(np.stack(ms), scores, (H, W), min_crys_size=2)
    got = mask_utils.postprocess_masks(np.stack(ms), scores, image, min_crys_size=2)
    assert len(got) == len(ref) and all(g.dtype == np.uint8 and np.array_equal(g, r) for g, r in zip(got, ref))
    assert mask_utils.postprocess_masks(np.stack(ms), np.zeros(len(ms), np.float32), image) == []
    for small in (True, False):
        ref = morphology.postprocess_masks_universal(ms, (H, W), small)
        got = inf.postprocess_masks_universal(np.stack(ms), scores, image, 0, small)
        assert len(got) == len(ref) and all(g.dtype == bool and np.array_equal(g, r) for g, r in zip(got, ref))
    ref = morphology.process_masks_parallel([m.astype(np.uint8) for m in ms])
    got = inf.process_masks_parallel([m.astype(np.uint8) for m in ms])
    assert all(np.array_equal(g, r) for g, r in zip(got, ref))


def test_dedup_and_spatial_mirrors(cuda_device):
    H, W = 200, 200
    ms, rng = _masks(3, 30, H, W, dup=20)
    scores = [np.float32(s) for s in syn.distinct_scores(rng, len(ms))]
    classes = [int(c) for c in (rng.random(len(ms)) < 0.5)]
    m, s, c = inf.deduplicate_masks_smart(ms, scores, classes, iou_threshold=0.4)
    rm, rs, rc, idx = dedup.deduplicate_masks_smart(ms, scores, classes, iou_threshold=0.4, return_indices=True)
    assert [id(x) for x in m] == [id(ms[i]) for i in idx] and s == rs and c == rc
    assert inf.deduplicate_masks_smart([], [], []) == ([], [], [])
    rules = syn.POLYHIPES_RULES
    gm, gs, gc, rem = sc.filter_by_overlap_rules(ms, scores, classes, rules['overlap_rules'])
    assert rem == ospatial.filter_by_overlap_rules(ms, scores, classes, rules['overlap_rules'])
    gm, gs, gc, rem = sc.filter_by_containment_rules(ms, scores, classes, rules['containment_rules'], 0.95)
    assert rem == ospatial.filter_by_containment_rules(ms, scores, classes, rules['containment_rules'], 0.95)
    gm, gs, gc = sc.apply_spatial_constraints(ms, scores, classes, rules=rules)
    _, _, _, idx = ospatial.apply_spatial_constraints(ms, scores, classes, rules)
    assert [id(x) for x in gm] == [id(ms[i]) for i in idx]
    sc.set_constraint_loader(lambda name: rules if name == "polyhipes_tommy" else None)
    try:
        assert len(sc.apply_spatial_constraints(ms, scores, classes, dataset_name="polyhipes_tommy")[0]) == len(idx)
        assert len(sc.apply_spatial_constraints(ms, scores, classes, dataset_name="other")[0]) == len(ms)
    finally:
        sc.set_constraint_loader(None)


def test_calculate_measurements_and_csv(cuda_device, tmp_path):
    H, W = 256, 300
    ms, rng = _masks(4, 20, H, W, dup=0)
    two = np.zeros((H, W), bool); two[20:40, 20:50] = True; two[60:64, 100:104] = True; ms.append(two)       # two contours, one tiny
    image = rng.integers(0, 255, (H, W, 3), dtype=np.uint8)
    # calculate_measurements on single contours: values AND numpy scalar types (Q14)
    for m in ms[:6]:
        binary = m.astype(np.uint8) * 255
        c = cv2.findContours(binary, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)[0][0]
        got = measurements.calculate_measurements(c, binary, um_pix=0.5, pixelsPerMetric=1)
        ref = omeasure.calculate_measurements(c, um_pix=0.5, pixelsPerMetric=1)
        assert list(got) == measurements.KEY_ORDER
        for k in omeasure.MEASUREMENT_KEYS:
            np.testing.assert_allclose(float(got[k]), float(ref[k]), rtol=1e-5, atol=0, err_msg=k)
            assert type(got[k]) is type(ref[k]), (k, type(got[k]), type(ref[k]))
    # contrast percentiles
    binary = ms[0].astype(np.uint8) * 255
    c = cv2.findContours(binary, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)[0][0]
    got = measurements.calculate_measurements(c, binary, um_pix=0.5, original_image=image, measure_contrast_distribution=True)
    gray = cv2.cvtColor(image, cv2.COLOR_BGR2GRAY)
    ref = omeasure.calculate_measurements(c, um_pix=0.5, gray=gray, mask=binary)
    for k in ("contrast_d10", "contrast_d50", "contrast_d90"):
        np.testing.assert_allclose(got[k], ref[k], rtol=1e-9, err_msg=k)
    # the masked grey histograms themselves: exact integer counts (cv2's 15-bit fixed-point BGR2GRAY)
    import torch
    from deepemia_b200 import engine
    iset = engine.from_masks(torch.as_tensor(np.stack(ms[:8]).astype(np.uint8), device=cuda_device))
    hist = engine.gray_hist(iset, image).cpu().numpy()
    for i in range(8):
        assert np.array_equal(hist[i], np.bincount(gray[ms[i]], minlength=256)), f"histogram of instance {i}"
    assert np.array_equal(engine.image_gray_hist(image).cpu().numpy(), np.bincount(gray.reshape(-1), minlength=256))
    # measurement loop rows
    classes = [int(v) for v in rng.integers(0, 2, len(ms))]
    rows = inf.measure_masks(ms, classes, image.shape, 0.5, "imgA.tif", "500", class_names=["pore", "throat"])
    ref_rows = omeasure.measure_masks(ms, classes, (H, W), 0.5, test_img="imgA.tif", class_names=["pore", "throat"], psum="500")
    assert len(rows) == len(ref_rows)
    for r, q in zip(rows, ref_rows):
        assert r[:3] == q[:3] and r[15:] == q[15:]
        np.testing.assert_allclose([float(v) for v in r[3:15]], [float(v) for v in q[3:15]], rtol=1e-5, atol=0)


def test_run_inference_writes_reference_csv_schema(cuda_device, tmp_path):
    imgs = [("a.png", np.random.default_rng(7).integers(0, 255, (160, 200, 3), dtype=np.uint8)),
            ("b.png", np.random.default_rng(8).integers(0, 255, (150, 180, 3), dtype=np.uint8))]
    pred = syn.FakeHeadPredictor(base_seed=9, n=24)
    out = inf.run_inference("polyhipes_tommy", str(tmp_path), images=imgs, predictors=[pred], thing_classes=["pore", "throat"],
                            small_classes={1}, spatial_rules=syn.POLYHIPES_RULES, tile_size=96, overlap_ratio=0.25,
                            scale_bar_fn=lambda im: ("500", 0.5))
    assert out is None
    rows = list(csv.reader(open(os.path.join(tmp_path, "measurements_results.csv"))))
    assert rows[0] == omeasure.CSV_HEADER and len(rows) > 1 and all(len(r) == 20 for r in rows)
    assert {r[19] for r in rows[1:]} <= {"a.png", "b.png"} and all(r[18] == "500" for r in rows[1:])
    rle = list(csv.reader(open(os.path.join(tmp_path, "R50_flip_results.csv"))))
    assert rle[0] == ["ImageId", "EncodedPixels"] and len(rle) > 1
    # every RLE row decodes to a mask whose measurement rows exist
    ids = {r[0].rsplit("_", 1)[0] for r in rows[1:]}
    assert ids <= {"a.png", "b.png"}
    legend = open(os.path.join(tmp_path, "class_color_legend.txt")).read().splitlines()
    assert legend[2:] == ["Class 0 (pore): RGB(0, 255, 0)", "Class 1 (throat): RGB(0, 0, 255)"]


def test_adaptive_confidence_threshold(cuda_device):
    """calculate_image_quality_score / adaptive_confidence_threshold / get_confidence_threshold (inference.py:256-362)."""
    rng = np.random.default_rng(5)
    for lo, hi in ((0, 40), (60, 200), (0, 256), (200, 256)):
        img = rng.integers(lo, hi, (301, 417, 3), dtype=np.uint8)
        gray = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
        ref = np.clip(0.4 * (np.mean(gray) / 255.0) + 0.6 * (np.std(gray) / 128.0), 0.0, 1.0)
        got = inf.calculate_image_quality_score(img)
        np.testing.assert_allclose(got, ref, rtol=1e-12)
        assert inf.calculate_image_quality_score(gray) == inf.calculate_image_quality_score(np.ascontiguousarray(gray))
        want = 0.5 * (0.7 if ref < 0.3 else 0.85 if ref < 0.5 else 1.0)
        assert inf.get_confidence_threshold(img, 0, {1}) == want
        assert inf.get_confidence_threshold(img, 1, {1}, {"class_1": {"confidence_threshold": 0.2}}, "manual") == 0.2
