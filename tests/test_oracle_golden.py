"""CPU: the oracle restatement (oracle/) against the golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py, run in the build container where /root/reference exists)."""
import os

import cv2
import numpy as np
import pytest

from oracle import d2_paste, dedup, measure, morphology, spatial, tiles
from deepemia_b200 import synthetic as syn

G = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    return np.load(os.path.join(G, name), allow_pickle=False)


def test_measurements_match_reference_golden():
    g = _load("measure_golden.npz")
    H, W = int(g["H"]), int(g["W"])
    ps, pp = g["poly_start"], g["poly_pts"]
    k = 0
    for i in range(len(ps) - 1):
        m = np.zeros((H, W), np.uint8)
        cv2.fillPoly(m, [pp[ps[i]:ps[i + 1]]], 255)
        um = [0.5, 1.0, 0.123][i % 3]
        for c in cv2.findContours(m, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)[0]:
            if cv2.contourArea(c) < 5:
                continue
            assert np.array_equal(c[:, 0, :], g["verts"][g["vstart"][k]:g["vstart"][k + 1]])
            out = measure.calculate_measurements(c, um_pix=um)
            vals = np.array([float(out[key]) for key in measure.MEASUREMENT_KEYS])
            assert np.array_equal(vals, g["vals"][k]), (i, vals, g["vals"][k])
            f32 = [isinstance(out[key], np.float32) for key in measure.MEASUREMENT_KEYS]
            assert f32 == list(g["f32_flags"])
            k += 1
    assert k == len(g["vals"])


def _group(g, k):
    H, W = int(g["H"]), int(g["W"])
    probs = g[f"probs{k}"].astype(np.float32)
    masks = d2_paste.paste_masks_in_image(probs, g[f"boxes{k}"], (H, W))
    return [m for m in masks], [np.float32(s) for s in g[f"scores{k}"]], [int(c) for c in g[f"classes{k}"]]


def test_dedup_and_spatial_match_reference_golden():
    g = _load("dedup_golden.npz")
    rules = syn.POLYHIPES_RULES
    for k in range(3):
        ml, sl, cl = _group(g, k)
        for thr in (0.4, 0.7):
            _, _, _, idx = dedup.deduplicate_masks_smart(ml, sl, cl, iou_threshold=thr, return_indices=True)
            assert idx == g[f"smart{k}_{int(thr * 10)}"].tolist()
        assert sorted(spatial.filter_by_overlap_rules(ml, sl, cl, rules['overlap_rules'])) == g[f"overlap_removed{k}"].tolist()
        assert sorted(spatial.filter_by_containment_rules(ml, sl, cl, rules['containment_rules'], 0.95)) == g[f"contain_removed{k}"].tolist()
        assert sorted(spatial.filter_by_containment_rules(ml, sl, cl, {1: 0}, 0.5)) == g[f"contain50_removed{k}"].tolist()
        _, _, _, kept = dedup.greedy_inorder_dedup(ml, sl, 0, 0.5)
        assert kept == g[f"inorder{k}"].tolist()
        pi = np.array([[dedup.iou(ml[a], ml[b]), dedup.calculate_iou(ml[a], ml[b]), spatial.calculate_iou(ml[a], ml[b]),
                        spatial.calculate_containment(ml[a], ml[b])] for a in range(12) for b in range(12)])
        assert np.array_equal(pi, g[f"pair_iou{k}"])


def test_dedup_quirk_kats():
    g = _load("dedup_golden.npz")

    def disc(x, y, r=8):
        m = np.zeros((128, 128), np.uint8); cv2.circle(m, (x, y), r, 1, -1); return m.astype(bool)
    kat = [([disc(20, 90), disc(20, 90)], [0.9, 0.8]), ([disc(90, 20), disc(90, 20)], [0.9, 0.8]),
           ([disc(64, 64)] * 3, [0.7, 0.8, 0.9]), ([disc(64, 64)] * 3, [0.9, 0.8, 0.7])]
    expected = [[0, 1], [0], [2, 1], [0]]
    for k, (ms, sc) in enumerate(kat):
        _, _, _, idx = dedup.deduplicate_masks_smart(ms, [np.float32(s) for s in sc], [0] * len(ms), 0.4, return_indices=True)
        assert idx == g[f"quirk{k}"].tolist() == expected[k]
    line = np.zeros((128, 128), bool); line[60, 10:100] = True
    assert dedup.deduplicate_masks_smart([line], [np.float32(0.9)], [0], 0.4, return_indices=True)[3] == g["thin_line_kept"].tolist() == []


def test_paste_restatement_pinned():
    g = _load("paste_golden.npz")
    H, W = int(g["H"]), int(g["W"])
    probs = g["probs"].astype(np.float32)
    for tag, (sx, sy) in {"a": (1.0, 1.0), "b": (1.28, 0.77)}.items():
        masks, _, _, _ = d2_paste.predictor_instances(probs, g["boxes"], np.ones(len(probs), np.float32), np.zeros(len(probs), int), sx, sy, H, W)
        _, keep = d2_paste.detector_postprocess_boxes(g["boxes"], sx, sy, H, W)
        assert np.array_equal(keep, g[f"keep_{tag}"])
        assert np.array_equal(np.packbits(masks, axis=-1, bitorder="little"), g[f"bits_{tag}"])


def test_misc_golden():
    g = _load("misc_golden.npz")
    H, W = int(g["H"]), int(g["W"])
    masks = np.unpackbits(g["masks"], axis=-1, bitorder="little")[:, :, :W].astype(np.uint8)
    for i, m in enumerate(masks):
        assert np.array_equal(np.array(morphology.rle_encoding(m), np.int64), g["rle"][g["rle_start"][i]:g["rle_start"][i + 1]])
        assert tiles.is_edge_mask(m.astype(bool), H, 0.25) == bool(g["is_edge"][i])
    for small in (True, False):
        tag = 'small' if small else 'large'
        ref = np.unpackbits(g[f"universal_{tag}"], axis=-1, bitorder="little")[:, :, :W].astype(bool)
        for i, m in enumerate(masks):
            r = morphology.postprocess_masks_universal([m.astype(bool)], (H, W), small)
            assert len(r) == int(g[f"universal_{tag}_kept"][i])
            if r:
                assert np.array_equal(r[0], ref[i])
    pm = morphology.postprocess_masks(masks.astype(bool), np.linspace(0.9, 0.6, len(masks)).astype(np.float32), (H, W), 5)
    assert np.array_equal(np.stack(pm).astype(bool), np.unpackbits(g["postprocess_masks"], axis=-1, bitorder="little")[:, :, :W].astype(bool))
    pp = morphology.process_masks_parallel(pm)
    assert np.array_equal(np.stack(pp).astype(bool), np.unpackbits(g["process_masks_parallel"], axis=-1, bitorder="little")[:, :, :W].astype(bool))
    xy = [(x, y) for _, x, y in tiles.generate_tiles_with_overlap(np.zeros((700, 900, 3), np.uint8), 256, 0.125)]
    assert np.array_equal(np.array(xy, np.int32), g["tiles_xy"])
    assert len(tiles.tile_origins(8192, 8192, 1024, 0.125)) == int(g["tiles_8192"]) == 100


# ---- oracle/flows.py (the CPU restatement of the FLOWS used by bench.py's CPU legs and the batched-flow parity tests) against
# the golden vectors produced by the UNMODIFIED reference functions (tests/golden/make_golden_flows.py) ------------------------
def _flow_golden():
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "golden"))
    import flow_cases
    return flow_cases, np.load(os.path.join(here, "golden", "flows_golden.npz"), allow_pickle=False)


def _instances(pred, image):
    from oracle import flows
    probs, boxes, scores, classes, (in_h, in_w) = pred.raw_heads(image)
    H, W = image.shape[:2]
    return flows.heads_to_instances(probs, boxes, scores, classes, H, W, W / in_w, H / in_h)


def _check_flow(gold, name, masks, scores):
    n = len(gold[f"{name}/scores"])
    assert len(masks) == n
    if n:
        h, w = (int(v) for v in gold[f"{name}/shape"])
        ref = np.unpackbits(gold[f"{name}/bits"], axis=1)[:, :h * w].reshape(n, h, w).astype(bool)
        for i in range(n):
            assert np.array_equal(np.asarray(masks[i]) != 0, ref[i]), (name, i)
            assert float(scores[i]) == float(gold[f"{name}/scores"][i])


def test_oracle_flows_match_reference_golden():
    import cv2
    from deepemia_b200 import synthetic as syn
    from oracle import flows, tiles
    fc, gold = _flow_golden()
    for name, case in fc.CASES.items():
        if case.get("ensemble") or case["fn"] not in ("run_class_specific_inference", "tile_based_inference_pipeline"):
            continue
        image = fc.make_image(case["image_seed"], *case["shape"])
        pred = syn.FakeHeadPredictor(**case["predictors"][0])
        kw = case["kwargs"]
        if case["fn"] == "run_class_specific_inference":
            t, small = case["args"]
            ms = (kw.get("class_specific_settings") or {}).get(f"class_{t}", {}).get("min_size")
            m, s, c = flows.run_class_specific_inference(_instances(pred, image), t, small, kw.get("confidence_threshold", 0.3),
                                                         kw.get("iou_threshold", 0.7), ms, case.get("parallel", True))
        else:
            t, small, conf = case["args"]
            ts, ov, up = kw["tile_size"], kw["overlap_ratio"], kw["upscale_factor"]
            tl = tiles.generate_tiles_with_overlap(image, ts, ov)
            tinst = [_instances(pred, cv2.resize(ti, (int(ts * up), int(ts * up)), interpolation=cv2.INTER_LINEAR)) for ti, _, _ in tl]
            m, s, c = flows.tile_based_inference_pipeline(_instances(pred, image), tinst, [(x, y) for _, x, y in tl], image.shape[:2], t,
                                                          small, conf, ts, ov, kw.get("iou_threshold", 0.7),
                                                          kw.get("edge_filter_enabled", True))
        _check_flow(gold, name, m, s)


def _scalebar_golden():
    import json
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "golden"))
    import scalebar_cases
    g = np.load(os.path.join(here, "golden", "scalebar_golden.npz"), allow_pickle=False)
    return scalebar_cases, g, json.loads(str(g["config_json"]))


def test_oracle_scalebar_matches_reference_golden():
    """oracle.scalebar (cv2 calls + the reference's scalar logic) == the unmodified detect_scale_bar on every case."""
    from oracle import scalebar as osb
    cases, g, cfg = _scalebar_golden()
    positives = 0
    for name in cases.CASES:
        image, ocr, kw = cases.build(name, cfg["scale_bar_rois"]["default"])
        roi = kw.pop("roi_config", None) or cfg["scale_bar_rois"]["default"]
        kw.pop("dataset_name", None)
        psum, um = osb.detect_scale_bar(image, ocr, roi, thresholds=cfg["scalebar_thresholds"], **kw)
        assert psum == str(g[name + "/psum"]), name
        assert float(um) == float(g[name + "/um_pix"]), name
        positives += psum != "0"
    assert positives >= 10


def test_scalebar_host_logic_matches_reference_golden():
    """The host half of the mirror (filters, merge_collinear_segments, selection) fed with OpenCV's lines / line sums reproduces the
    reference result of every golden case (the device half is checked against OpenCV in tests/test_gpu_scalebar.py)."""
    from oracle import scalebar as osb
    from deepemia_b200.utils import scalebar_ocr as sb
    cases, g, cfg = _scalebar_golden()
    th = cfg["scalebar_thresholds"]
    sb.set_config_provider(lambda name: cfg)
    try:
        for name in cases.CASES:
            image, ocr, kw = cases.build(name, cfg["scale_bar_rois"]["default"])
            roi = kw.pop("roi_config", None) or sb.get_scalebar_roi_for_dataset(kw.get("dataset_name"))
            it, pt, gap, mll, emf = sb._thresholds(kw.get("dataset_name"), kw.get("intensity_threshold", 200), kw.get("proximity_threshold", 50))
            assert (gap, mll, emf) == (th["merge_gap"], th["min_line_length"], th["edge_margin_factor"])
            d = {}
            osb.detect_scale_bar(image, ocr, roi, thresholds=th, details=d)
            psum, centre, _ = sb._first_number(ocr)
            longest, length = None, 0
            if centre is not None and d["lines"] is not None:
                lines = d["lines"][:, 0, :]
                sums = []
                for x1, y1, x2, y2 in lines:
                    m = np.zeros_like(d["gray"])
                    cv2.line(m, (int(x1), int(y1)), (int(x2), int(y2)), 255, 2)
                    sums.append((int(d["gray"][m > 0].sum()), int((m > 0).sum())))
                longest, length, _ = sb.select_scale_line(lines, sums, centre, d["gray"].shape[1], d["gray"].shape[0], it, pt, gap, mll, emf)
            got = (psum, float(psum) / length) if longest else ("0", 1)
            assert got[0] == str(g[name + "/psum"]) and float(got[1]) == float(g[name + "/um_pix"]), name
    finally:
        sb.set_config_provider(None)
