"""GPU parity of the round-2 additions outside the flows: cv2.moments, colour sums / wavelength, the visualisation overlay
(cv2.addWeighted + cv2.drawContours), the torchvision head adapter + mask-size heuristic, and run_inference called the way the
reference's main.py:480-488 calls it (providers)."""
import csv
import os

import cv2
import numpy as np
import pytest
import torch

from deepemia_b200 import adapters, engine, synthetic as syn
from deepemia_b200.functions import inference as inf
from deepemia_b200.utils import _bridge, measurements as M, visualize

pytestmark = pytest.mark.gpu


def _masks(seed, n, H, W):
    rng = np.random.default_rng(seed)
    return syn.masks_from_polys(syn.particle_field(rng, n, H, W, rmin=4, rmax=25, margin=28), H, W)


def test_moments_match_opencv(cuda_device):
    H, W = 300, 420
    masks = _masks(3, 60, H, W) + [np.zeros((H, W), np.uint8)]
    iset = _bridge.upload(masks)
    got = engine.moments(iset).cpu().numpy()
    for i, m in enumerate(masks):
        ref = cv2.moments(np.ascontiguousarray(m))
        want = np.array([ref[k] for k in engine.MOMENT_NAMES])
        assert np.array_equal(got[i, :13], want[:13]), i                           # raw + second-order central: bit-exact
        assert np.allclose(got[i, 13:17], want[13:17], rtol=1e-9, atol=1e-12 * max(want[0], 1) * W ** 3)
        assert np.allclose(got[i, 17:], want[17:], rtol=1e-6, atol=1e-8 + 1e-13 * W ** 3 / max(want[0], 1.0) ** 1.5)
    m01 = engine.moments01(iset).cpu().numpy()
    assert np.array_equal(m01, got[:, :3].astype(np.int64))


def test_color_sums_and_wavelength(cuda_device):
    H, W = 200, 260
    masks = _masks(4, 30, H, W)
    img = np.random.default_rng(5).integers(0, 256, (H, W, 3), dtype=np.uint8)
    iset = _bridge.upload(masks)
    got = engine.color_sums(iset, img).cpu().numpy()
    wl = M.instance_wavelengths(iset, img)
    for i, m in enumerate(masks):
        sel = img[m.astype(bool)]
        assert got[i].tolist() == [int(sel[:, 0].sum()), int(sel[:, 1].sum()), int(sel[:, 2].sum()), int(m.sum())]
        b, g, r = (sel[:, k].sum() / m.sum() for k in range(3))
        assert wl[i] == M.rgb_to_wavelength(r, g, b)


def test_overlay_matches_opencv(cuda_device):
    """Blends and outlines of overlapping masks in list order == the reference's loop (inference.py:1080-1100), bit for bit."""
    H, W = 240, 300
    masks = _masks(6, 40, H, W)
    classes = [int(c) for c in np.random.default_rng(7).integers(0, 10, len(masks))]
    img = np.random.default_rng(8).integers(0, 256, (H, W, 3), dtype=np.uint8)
    ref = img.copy()
    for mask, cls in zip(masks, classes):
        color = visualize.CLASS_COLORS_BGR[cls % len(visualize.CLASS_COLORS_BGR)]
        colored = np.zeros_like(ref)
        colored[mask.astype(bool)] = color
        ref = cv2.addWeighted(ref, 1.0, colored, 0.5, 0)
        contours, _ = cv2.findContours(mask.astype(np.uint8), cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
        cv2.drawContours(ref, contours, -1, color, 1)
    iset = _bridge.upload(masks)
    engine.trace(iset)
    got = visualize.overlay(img, iset, classes).cpu().numpy()
    assert np.array_equal(got, ref)
    # a subset, in a different order
    order = [7, 3, 11, 0]
    ref2 = img.copy()
    for k in order:
        color = visualize.CLASS_COLORS_BGR[classes[k] % 8]
        colored = np.zeros_like(ref2); colored[masks[k].astype(bool)] = color
        ref2 = cv2.addWeighted(ref2, 1.0, colored, 0.5, 0)
        cs, _ = cv2.findContours(masks[k].astype(np.uint8), cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
        cv2.drawContours(ref2, cs, -1, color, 1)
    assert np.array_equal(visualize.overlay(img, iset, classes, order=order).cpu().numpy(), ref2)
    vis = visualize.render_predictions(img, iset, classes, ["pore", "throat"])
    assert vis.shape == img.shape and not np.array_equal(vis, ref)              # labels on top


def test_torchvision_adapter_and_size_heuristic(cuda_device):
    """TorchvisionHeadAdapter == the model's own forward up to (and excluding) torchvision's paste; the size heuristic
    (calculate_average_mask_sizes / determine_small_classes, inference.py:1626-1736) from K1's areas == from the oracle's masks."""
    pytest.importorskip("torchvision")
    import torchvision.models.detection.roi_heads as rh
    from torchvision.models.detection import maskrcnn_resnet50_fpn
    from oracle import d2_paste
    torch.manual_seed(0)
    model = maskrcnn_resnet50_fpn(weights=None, weights_backbone=None, box_score_thresh=0.0, num_classes=3, min_size=512, max_size=683).eval().to(cuda_device)
    ad = adapters.TorchvisionHeadAdapter(model)
    image = np.random.default_rng(1).integers(0, 256, (384, 512, 3), dtype=np.uint8)
    ho = ad.heads(image)
    n = ho.probs.shape[0]
    assert n > 0 and ho.probs.shape[1:] == (28, 28) and ho.boxes.shape == (n, 4) and ho.classes.min() >= 0
    # the model's own forward (which rescales the boxes to the original image) sees the same detections
    captured = {}
    orig = rh.maskrcnn_inference

    def spy(x, labels):
        out = orig(x, labels)
        captured["probs"] = out[0][:, 0].detach().float()
        return out
    rh.maskrcnn_inference = spy
    try:
        x = torch.as_tensor(np.ascontiguousarray(image[:, :, ::-1].transpose(2, 0, 1))).to(cuda_device).float() / 255.0
        with torch.no_grad():
            det = model([x])[0]
    finally:
        rh.maskrcnn_inference = orig
    assert torch.equal(captured["probs"], ho.probs.float()) and torch.equal(det["scores"], ho.scores)
    in_h, in_w = ho.input_size
    sx, sy = image.shape[1] / in_w, image.shape[0] / in_h
    np.testing.assert_allclose(det["boxes"].cpu().numpy(), ho.boxes.cpu().numpy() * np.array([sx, sy, sx, sy], np.float32), rtol=1e-5, atol=1e-3)
    # heuristic: scores of a random-init head are low, so use a stand-in predictor with confident scores
    pred = syn.FakeHeadPredictor(base_seed=3, n=40, duplicate_frac=0.0)
    imgs = [np.random.default_rng(20 + k).integers(0, 255, (200, 240, 3), dtype=np.uint8) for k in range(3)]
    got = adapters.calculate_average_mask_sizes([pred], imgs)
    want = {}
    for im in imgs:
        probs, boxes, scores, classes, (ih, iw) = pred.raw_heads(im)
        masks, s, c, _ = d2_paste.predictor_instances(probs, boxes, scores, classes, 240 / iw, 200 / ih, 200, 240)
        for mk, sc_, cl in zip(masks, s, c):
            if sc_ >= 0.7:
                want.setdefault(int(cl), []).append(np.sum(mk))
    want = {k: np.mean(v) for k, v in want.items()}
    assert got.keys() == want.keys() and all(got[k] == want[k] for k in want)
    assert adapters.determine_small_classes(got) == {c for c, v in want.items() if v <= np.percentile(list(want.values()), 50)}


def test_run_inference_called_like_main_py(cuda_device, tmp_path):
    """run_inference(dataset_name, output_dir, visualize=..., threshold=..., draw_id=..., dataset_format=..., draw_scalebar=...)
    (main.py:480-488) with the reference's subsystems behind providers == the explicit-keyword form."""
    from deepemia_b200.utils import spatial_constraints as sc
    imgs = [("a.png", np.random.default_rng(7).integers(0, 255, (160, 200, 3), dtype=np.uint8)),
            ("b.tif", np.random.default_rng(8).integers(0, 255, (150, 180, 3), dtype=np.uint8))]
    pred = syn.FakeHeadPredictor(base_seed=9, n=24)
    cfg = {"inference_settings": {"confidence_mode": "manual", "tile_settings": {"tile_size": 96, "overlap_ratio": 0.25},
                                  "class_specific_settings": {"class_0": {"confidence_threshold": 0.3}, "class_1": {"confidence_threshold": 0.2}}}}
    calls = {}

    def predictors(name, threshold, thing_classes):
        calls["predictors"] = (name, threshold, list(thing_classes))
        return [pred]
    inf.clear_providers()
    with pytest.raises(FileNotFoundError):
        inf.run_inference("polyhipes_tommy", str(tmp_path / "none"))           # nothing registered: the reference's own error type
    inf.set_providers(config=lambda name: cfg, thing_classes=lambda name: ["pore", "throat"], predictors=predictors,
                      images=lambda name: imgs, scale_bar=lambda im, roi, name: ("500", 0.5))
    sc.set_constraint_loader(lambda name: syn.POLYHIPES_RULES)
    try:
        d1 = tmp_path / "providers"
        assert inf.run_inference("polyhipes_tommy", str(d1), True, 0.65, False, "json", False) is None
        assert calls["predictors"] == ("polyhipes_tommy", 0.65, ["pore", "throat"])
    finally:
        inf.clear_providers()
        sc.set_constraint_loader(None)
    d2 = tmp_path / "explicit"
    small = adapters.determine_small_classes(adapters.calculate_average_mask_sizes([pred], [im for _, im in imgs]))
    inf.run_inference("polyhipes_tommy", str(d2), visualize=False, images=imgs, predictors=[pred], thing_classes=["pore", "throat"],
                      small_classes=small, spatial_rules=syn.POLYHIPES_RULES, tile_size=96, overlap_ratio=0.25, confidence_mode="manual",
                      class_specific_settings=cfg["inference_settings"]["class_specific_settings"], scale_bar_fn=lambda im: ("500", 0.5))
    for f in ("measurements_results.csv", "R50_flip_results.csv", "class_color_legend.txt"):
        assert open(d1 / f).read() == open(d2 / f).read(), f
    rows = list(csv.reader(open(d1 / "measurements_results.csv")))
    assert rows[0] == inf.CSV_HEADER and len(rows) > 3
    assert os.path.exists(d1 / "a.png_predictions.png") and os.path.exists(d1 / "b.tif_predictions.png")
    # the RLE rows decode to the masks whose measurements were written: areas agree with K1's popcounts through rle_encoding
    rle = list(csv.reader(open(d1 / "R50_flip_results.csv")))
    assert rle[0] == ["ImageId", "EncodedPixels"] and {r[0] for r in rle[1:]} <= {"a", "b"}
    assert all(len(r[1].split()) % 2 == 0 and len(r[1]) > 0 for r in rle[1:])
