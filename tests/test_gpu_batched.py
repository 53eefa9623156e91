"""GPU parity of the BATCHED flows (deepemia_b200.batched) — one launch per stage over all tiles / images / classes, no host
read-back inside a flow — against (a) the golden vectors of the UNMODIFIED reference functions (tests/golden/flows_golden.npz:
run_class_specific_inference, tile_based_inference_pipeline of src/functions/inference.py) and (b) the per-image mirrors, which
are themselves pinned by the same golden file (tests/test_gpu_flows.py).  Masks, kept sets and their order are bit-exact."""
import os
import sys

import cv2
import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import flow_cases  # noqa: E402

from deepemia_b200 import batched, engine, synthetic as syn  # noqa: E402
from deepemia_b200.functions import inference as inf  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(HERE, "golden", "flows_golden.npz"), allow_pickle=False)


def head_batch(pred, images, dev):
    """HeadBatch of `pred` on a list of same-size images (one unit per image)."""
    parts = [pred.raw_heads(im) for im in images]
    H, W = images[0].shape[:2]
    in_h, in_w = parts[0][4]
    off = np.concatenate([[0], np.cumsum([len(p[2]) for p in parts])]).astype(np.int64)
    cat = lambda k, dt: torch.as_tensor(np.ascontiguousarray(np.concatenate([p[k] for p in parts]).astype(dt)), device=dev)
    return batched.HeadBatch(cat(0, np.float32), cat(1, np.float32), cat(2, np.float32), cat(3, np.int32), off, H, W,
                             scale_x=float(W) / in_w, scale_y=float(H) / in_h)


def run_flow(fn):
    """Run a sync-free flow until no capacity guard trips (at most a few re-runs while the arena grows)."""
    arena = engine.Arena(torch.device("cuda", torch.cuda.current_device()))
    for _ in range(8):
        arena.begin()
        res = fn(arena)
        if arena.finish():
            return res, arena
    raise AssertionError("the arena did not converge")


def check_golden(name, iset, ids, scores):
    g = {k: GOLD[f"{name}/{k}"] for k in ("bits", "scores", "classes", "shape")}
    n = len(g["scores"])
    assert len(ids) == n, (len(ids), n)
    if n == 0:
        return
    h, w = (int(v) for v in g["shape"])
    ref = np.unpackbits(g["bits"], axis=1)[:, :h * w].reshape(n, h, w).astype(bool)
    got = engine.unpack_masks(iset, ids).cpu().numpy() != 0
    assert got.shape == ref.shape
    for i in range(n):
        assert np.array_equal(got[i], ref[i]), f"mask {i} differs"
        assert float(scores[ids[i]]) == float(g["scores"][i]), f"score {i}"


CS_CASES = [k for k, v in flow_cases.CASES.items() if v["fn"] == "run_class_specific_inference" and not v.get("ensemble")]


@pytest.mark.parametrize("name", CS_CASES)
def test_class_specific_batched_matches_reference(cuda_device, name):
    case = flow_cases.CASES[name]
    image = flow_cases.make_image(case["image_seed"], *case["shape"])
    pred = syn.FakeHeadPredictor(**case["predictors"][0])
    target, small = case["args"]
    kw = case["kwargs"]
    settings = kw.get("class_specific_settings") or {}
    p = batched.ClassParams(target, target in small, kw.get("confidence_threshold", 0.3), kw.get("iou_threshold", 0.7),
                            settings.get(f"class_{target}", {}).get("min_size"))
    hb = head_batch(pred, [image], cuda_device)
    old = batched.PARALLEL_MASK_PROCESSING
    batched.PARALLEL_MASK_PROCESSING = case.get("parallel", True)
    try:
        (post, kept), arena = run_flow(lambda a: batched.class_specific(hb, [p], a))
    finally:
        batched.PARALLEL_MASK_PROCESSING = old
    ids = kept.section(0).to_lists()[0]
    check_golden(name, post, ids, post.scores.cpu().numpy())


def tile_batches(pred, images, tile_size, overlap, upscale, dev):
    tiles_xy = None
    ups = []
    for im in images:
        tl = inf.generate_tiles_with_overlap(im, tile_size, overlap)
        tiles_xy = np.array([[x, y] for _, x, y in tl], np.int32)
        for t, _, _ in tl:
            ups.append(cv2.resize(t, (int(tile_size * upscale), int(tile_size * upscale)), interpolation=cv2.INTER_LINEAR))
    return head_batch(pred, images, dev), head_batch(pred, ups, dev), tiles_xy


@pytest.mark.parametrize("name", ["tile_pipeline", "tile_pipeline_no_edge_filter"])
def test_tile_pipeline_batched_matches_reference(cuda_device, name):
    case = flow_cases.CASES[name]
    image = flow_cases.make_image(case["image_seed"], *case["shape"])
    pred = syn.FakeHeadPredictor(**case["predictors"][0])
    target, small, conf = case["args"]
    kw = case["kwargs"]
    full_hb, tile_hb, xy = tile_batches(pred, [image], kw["tile_size"], kw["overlap_ratio"], kw["upscale_factor"], cuda_device)
    p = batched.ClassParams(target, target in small, conf, kw.get("iou_threshold", 0.7))
    res, arena = run_flow(lambda a: batched.tile_pipeline(full_hb, tile_hb, xy, image.shape[:2], kw["tile_size"], kw["overlap_ratio"], [p], a,
                                                          edge_filter_enabled=kw.get("edge_filter_enabled", True), measure=False))
    ids = res.per_class.to_lists()[0]
    check_golden(name, res.iset, ids, res.iset.scores.cpu().numpy())
    assert arena.aborts <= 3


def test_tile_pipeline_batch_of_images_and_classes(cuda_device):
    """B = 3 images x C = 2 classes through ONE batched run == the per-image mirror (infer_image), incl. the cross-class
    de-dup, the spatial rules and the morphometry records."""
    pred = syn.FakeHeadPredictor(base_seed=7, n=30, duplicate_frac=0.4)
    images = [flow_cases.make_image(70 + k, 192, 224) for k in range(3)]
    ts, ov, up = 96, 0.25, 2.0
    full_hb, tile_hb, xy = tile_batches(pred, images, ts, ov, up, cuda_device)
    params = [batched.ClassParams(0, False, 0.2, 0.7), batched.ClassParams(1, True, 0.15, 0.7)]
    rules = syn.POLYHIPES_RULES
    res, arena = run_flow(lambda a: batched.tile_pipeline(full_hb, tile_hb, xy, images[0].shape[:2], ts, ov, params, a, rules=rules, um_pix=0.5))
    res.meas.finalize()
    kept = res.kept.to_lists()
    rows = res.meas.rows_to_host()
    sc = res.iset.scores.cpu().numpy()
    settings = {"class_0": {"confidence_threshold": 0.2, "iou_threshold": 0.7}, "class_1": {"confidence_threshold": 0.15, "iou_threshold": 0.7}}
    total = 0
    for b, im in enumerate(images):
        masks, scores, classes = inf.infer_image([pred], im, 2, {1}, class_specific_settings=settings, confidence_mode='manual',
                                                 tile_size=ts, overlap_ratio=ov, upscale_factor=up, spatial_rules=rules, ensemble_enabled=False)
        assert len(masks) == len(kept[b]), (b, len(masks), len(kept[b]))
        if not masks:
            continue
        got = engine.unpack_masks(res.iset, kept[b]).cpu().numpy() != 0
        for i, m in enumerate(masks):
            assert np.array_equal(got[i], np.asarray(m) != 0), (b, i)
            assert float(scores[i]) == float(sc[kept[b][i]])
        # morphometry of the same masks through the mirror
        want = inf.measure_masks(masks, classes, im.shape, 0.5, "x", "0")
        have = [r for _, rr in rows[b] for r in rr if r[engine.REC_MEASURED] == 1.0]
        assert len(want) == len(have)
        for wr, hr in zip(want, have):
            np.testing.assert_allclose(np.array([float(v) for v in wr[3:15]]), hr[:12], rtol=1e-5, atol=0)
        total += len(masks)
    assert total > 20


def test_sparse_group_path_matches_fused(cuda_device):
    """The sparse K4 path (groups beyond 1024 slots) and the fused one-CTA-per-group kernels agree on every operation, on lists
    where both apply; and a 3 000-member list (sparse only) is idempotent under de-duplication."""
    H = W = 512
    probs, boxes, scores, classes = syn.synthetic_heads(99, 900, H, W, duplicate_frac=0.5, rmin=6, rmax=18, margin=20)
    t = [torch.as_tensor(a, device=cuda_device) for a in (probs, boxes, scores, classes)]
    iset = engine.paste(t[0], t[1], H, W, scores=t[2], classes=t[3])
    engine.trace(iset)
    lists = [list(range(0, 300)), list(range(300, 900)), [], list(range(100, 500))[::-1]]
    g = engine.groups_from_lists(lists, cuda_device)
    rules = syn.POLYHIPES_RULES
    out = {}
    for fused in (True, False):
        engine.FUSED_K4 = fused
        try:
            out[fused] = [engine.dedup_smart(iset, g, 0.4).to_lists(), engine.dedup_inorder(iset, g, 0.5).to_lists(),
                          engine.dedup_sorted(iset, g, 0.4).to_lists(),
                          engine.overlap_rules(iset, g, rules['overlap_rules']).to_lists(),
                          engine.containment_rules(iset, g, rules['containment_rules'], 0.95).to_lists()]
        finally:
            engine.FUSED_K4 = True
    for a, b in zip(out[True], out[False]):
        assert a == b
    assert sum(len(l) for l in out[True][0]) < 1300           # something was removed
    # sparse only (a 1 200-member list): against the oracle's deduplicate_masks_smart / iou() loops on the unpacked masks
    from oracle import dedup as odedup
    sub = engine.select(iset, list(range(400)))
    big = engine.concat([sub, sub, sub])
    sc = np.random.default_rng(5).permutation(big.n).astype(np.float32) / big.n
    big.scores = torch.as_tensor(sc, device=cuda_device)
    engine.trace(big)
    gb = engine.groups_from_lists([list(range(big.n))], cuda_device)
    masks = [m for m in (engine.unpack_masks(big).cpu().numpy() != 0)]
    cl = [int(c) for c in big.classes.cpu().numpy()]
    got = engine.dedup_smart(big, gb, 0.4).to_lists()[0]
    want = odedup.deduplicate_masks_smart(masks, [np.float32(v) for v in sc], cl, 0.4, return_indices=True)[3]
    assert got == [int(i) for i in want] and len(got) < big.n
    got = engine.dedup_inorder(big, gb, 0.5).to_lists()[0]
    want = odedup.greedy_inorder_dedup(masks, sc, 0, 0.5)[3]
    assert got == want
    # pathological: 400 copies of one mask in a 2 000-slot list overflow the edge list -> exact fallback
    one = engine.select(iset, [5] * 400 + list(range(900)) + list(range(700)))
    one.scores = torch.as_tensor(np.linspace(0.9, 0.1, one.n).astype(np.float32), device=cuda_device)
    engine.trace(one)
    go = engine.groups_from_lists([list(range(one.n))], cuda_device)
    kk = engine.dedup_sorted(one, go, 0.4).to_lists()[0]
    assert sum(1 for i in kk if i < 400) == 1 and kk[0] == 0
