"""GPU parity tests: the CUDA path (through the C ABI, libemia.so) against the CPU oracle on the same seeded inputs.
Bit-exact for masks, areas, bounding boxes, kept-instance sets, contour vertices, contourArea and arcLength;
<= 1e-5 relative for the float measurement columns (the tolerance BASELINE.json's north_star states)."""
import numpy as np
import pytest
import torch

from deepemia_b200 import engine, synthetic as syn
from oracle import d2_paste, dedup, measure, pipeline, spatial

pytestmark = pytest.mark.gpu
REL_TOL = 1e-5


@pytest.fixture(params=["fused", "staged"])
def k4_path(request):
    """Both K4 implementations: one CTA per group in shared memory (groups <= 1024) and the staged global-memory kernels."""
    old = engine.FUSED_K4
    engine.FUSED_K4 = request.param == "fused"
    yield request.param
    engine.FUSED_K4 = old


def _heads(seed, n, H, W, **kw):
    probs, boxes, scores, classes = syn.synthetic_heads(seed, n, H, W, **kw)
    return probs, boxes, scores, classes


def _dev(dev, *arrs):
    return [torch.as_tensor(np.ascontiguousarray(a), device=dev) for a in arrs]


def _unpack_frames(frames, W):
    """int32 [n,H,pw] bit frames -> bool [n,H,W]"""
    f = frames.cpu().numpy().view(np.uint32)
    bits = ((f[..., None] >> np.arange(32, dtype=np.uint32)) & 1).astype(bool)
    return bits.reshape(f.shape[0], f.shape[1], -1)[:, :, :W]


@pytest.mark.parametrize("variant", [0, 1, 2])
@pytest.mark.parametrize("shape", [(256, 320), (300, 333), (128, 1100)])
def test_paste_matches_detectron2_oracle(cuda_device, variant, shape):
    H, W = shape
    probs, boxes, scores, classes = _heads(11 + H, 48, H, W, rmin=4, rmax=25, margin=10)
    # edge cases: box hanging over every frame edge, degenerate boxes, a box covering the whole frame, tiny box
    extra = np.array([[-10, -10, 30, 25], [W - 20, H - 15, W + 30, H + 9], [5, 5, 5, 40], [7, 9, 30, 9],
                      [0, 0, W, H], [50.25, 60.5, 51.0, 61.25], [-50, -50, -10, -10], [W - 0.5, 3, W + 5, 30]], np.float32)
    eprobs = np.random.default_rng(5).random((len(extra), 28, 28)).astype(np.float16).astype(np.float32)
    probs = np.concatenate([probs, eprobs]); boxes = np.concatenate([boxes, extra])
    scores = syn.distinct_scores(np.random.default_rng(1), len(boxes)); classes = np.zeros(len(boxes), np.int32)
    for sx, sy in ((1.0, 1.0), (1.28, 0.77)):
        ref_masks, _, _, _ = d2_paste.predictor_instances(probs, boxes, scores, classes, sx, sy, H, W)
        _, keep = d2_paste.detector_postprocess_boxes(boxes, sx, sy, H, W)
        tp, tb = _dev(cuda_device, probs, boxes)
        iset = engine.paste(tp, tb, H, W, scale_x=sx, scale_y=sy, frames=True, variant=variant)
        torch.cuda.synchronize()
        assert np.array_equal(iset.valid.cpu().numpy(), keep)
        got = engine.unpack_masks(iset).cpu().numpy().astype(bool)
        fr = _unpack_frames(iset.frames, W)
        k = 0
        for i in range(len(boxes)):
            if not keep[i]:
                assert not got[i].any() and not fr[i].any()
                continue
            assert np.array_equal(got[i], ref_masks[k]), f"crop bits differ for instance {i}"
            assert np.array_equal(fr[i], ref_masks[k]), f"frame bits differ for instance {i}"
            assert int(iset.area[i]) == int(ref_masks[k].sum())
            bb = dedup.get_mask_bbox(ref_masks[k])
            assert tuple(iset.bbox[i].tolist()) == (tuple(int(v) for v in bb) if bb is not None else (-1, -1, -1, -1))
            k += 1


def test_paste_fp16_heads_equal_fp32(cuda_device):
    """AMP heads emit fp16 probabilities; K1 widens them exactly, so the masks equal those of the same values passed as fp32."""
    H, W = 300, 333 + 19
    probs, boxes, _, _ = _heads(77, 60, H, W, rmin=4, rmax=25, margin=10)
    assert np.array_equal(probs, probs.astype(np.float16).astype(np.float32))
    tp, tb = _dev(cuda_device, probs, boxes)
    a = engine.paste(tp, tb, H, W, frames=True)
    b = engine.paste(tp.to(torch.float16).contiguous(), tb, H, W, frames=True)
    assert torch.equal(a.crops[:a.total_crop_words], b.crops[:b.total_crop_words]) and torch.equal(a.frames, b.frames)
    assert torch.equal(a.area, b.area) and torch.equal(a.bbox, b.bbox)
    # frames wider than the variant-2 kernel handles: the halves are widened on the way in
    Hw, Ww = 40, 2500
    pw, bw, _, _ = _heads(78, 12, Hw, Ww, rmin=4, rmax=12, margin=14)
    tpw, tbw = _dev(cuda_device, pw, bw)
    c = engine.paste(tpw, tbw, Hw, Ww)
    d = engine.paste(tpw.to(torch.float16).contiguous(), tbw, Hw, Ww)
    assert torch.equal(c.crops[:c.total_crop_words], d.crops[:d.total_crop_words]) and torch.equal(c.area, d.area)


def test_exclusive_scan(cuda_device):
    rng = np.random.default_rng(0)
    for n in (0, 1, 5, 4096, 4097, 100_003, 1_022_339, 4_300_000):
        a = rng.integers(0, 1000, n + 1).astype(np.int64)
        t = torch.as_tensor(a, device=cuda_device)
        engine.exclusive_scan_(t)
        ref = np.concatenate([[0], np.cumsum(a[:n])])
        assert np.array_equal(t.cpu().numpy(), ref), n


def test_paste_frame_ring_and_crops_only(cuda_device):
    H, W = 256, 256
    probs, boxes, _, _ = _heads(3, 40, H, W, rmin=5, rmax=20, margin=10)
    tp, tb = _dev(cuda_device, probs, boxes)
    ref = d2_paste.predictor_instances(probs, boxes, np.ones(len(boxes), np.float32), np.zeros(len(boxes), np.int32), 1.0, 1.0, H, W)[0]
    assert len(ref) == len(boxes)
    ring = torch.full((len(probs), H, engine.pitch_words_for(W)), -1, dtype=torch.int32, device=cuda_device)
    iset = engine.paste(tp, tb, H, W, frames=ring)
    assert np.array_equal(_unpack_frames(ring, W), ref)
    # a ring smaller than n is a scratch arena (slot = i % slots, last writer undefined): only crops/bbox/area are defined
    small = torch.empty((8, H, engine.pitch_words_for(W)), dtype=torch.int32, device=cuda_device)
    iset3 = engine.paste(tp, tb, H, W, frames=small)
    assert torch.equal(iset.crops[:iset.total_crop_words], iset3.crops[:iset3.total_crop_words])
    iset2 = engine.paste(tp, tb, H, W, frames=None)
    assert torch.equal(iset.crops[:iset.total_crop_words], iset2.crops[:iset2.total_crop_words])
    assert torch.equal(iset.area, iset2.area) and torch.equal(iset.bbox, iset2.bbox)


def _check_records(iset, masks, classes, H, W, um):
    rows = measure.measure_masks(masks, classes, (H, W), um)
    rec = iset.records.cpu().numpy()
    inst = iset.rec_inst.cpu().numpy()
    sel = rec[:, engine.REC_MEASURED] == 1.0
    got_rows = rec[sel]
    got_inst = inst[sel]
    assert len(got_rows) == len(rows)
    exact = 0
    for r, g, gi in zip(rows, got_rows, got_inst):
        assert r[0] == f"img_{gi + 1}"
        ref_vals = np.array([float(v) for v in r[3:15]])
        if np.array_equal(ref_vals, g[:12]):
            exact += 1
        np.testing.assert_allclose(g[:12], ref_vals, rtol=REL_TOL, atol=0)
    return exact, len(rows)


@pytest.mark.parametrize("single_pass", [True, False])
def test_contours_and_measurements_match_opencv(cuda_device, single_pass):
    import cv2
    H, W = 512, 640
    rng = np.random.default_rng(1000)
    polys = syn.particle_field(rng, 120, H, W)
    masks = syn.masks_from_polys(polys, H, W)
    # adversarial shapes: line, single pixel, ring, two components, touching the frame, diagonal spur, empty
    adv = []
    m = np.zeros((H, W), np.uint8); m[100, 50:140] = 1; adv.append(m)
    m = np.zeros((H, W), np.uint8); m[7, 9] = 1; adv.append(m)
    m = np.zeros((H, W), np.uint8); cv2.circle(m, (200, 200), 30, 1, 3); adv.append(m)
    m = np.zeros((H, W), np.uint8); cv2.circle(m, (100, 300), 12, 1, -1); cv2.circle(m, (160, 310), 9, 1, -1); adv.append(m)
    m = np.zeros((H, W), np.uint8); m[0:30, 0:45] = 1; m[H - 20:, W - 33:] = 1; adv.append(m)
    m = np.zeros((H, W), np.uint8); cv2.circle(m, (400, 100), 15, 1, -1)
    for k in range(12): m[100 - 15 - k, 400 + 15 + k] = 1
    adv.append(m)
    adv.append(np.zeros((H, W), np.uint8))
    m = np.zeros((H, W), np.uint8); m[40:60, 31:33] = 1; m[40:42, 20:70] = 1; adv.append(m)      # crosses a word boundary
    masks = masks + adv
    classes = [0] * len(masks)
    t = torch.as_tensor(np.stack(masks), device=cuda_device)
    iset = engine.from_masks(t)
    engine.measure(iset, um_pix=0.5, single_pass=single_pass)
    torch.cuda.synchronize()
    # contour vertices, exact
    cont_off = iset.cont_off.cpu().numpy()
    rec = iset.records.cpu().numpy()
    got = engine.contours_to_host(iset)
    for i, mk in enumerate(masks):
        ref = cv2.findContours((mk > 0).astype(np.uint8) * 255, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)[0]
        assert len(got[i]) == len(ref), f"instance {i}: {len(got[i])} contours vs {len(ref)}"
        for j, c in enumerate(ref):
            assert np.array_equal(got[i][j], c[:, 0, :]), f"instance {i} contour {j}"
            r = rec[cont_off[i] + j]
            assert r[engine.REC_AREA] == cv2.contourArea(c) and r[engine.REC_PERIM] == cv2.arcLength(c, True)
            assert r[engine.REC_NVERT] == len(c)
    exact, total = _check_records(iset, masks, classes, H, W, 0.5)
    print(f"PARITY-COUNT contours[{single_pass}]: rows {total} bit-exact {exact}")
    assert total == 127 and exact >= 126, f"only {exact}/{total} rows bit-exact"       # measured on B200: 126 of 127
    # contours below the area gate are skipped before calculate_measurements (src/functions/inference.py:1176-1190): their records
    # carry area / perimeter / vertex count (checked above) and nothing else
    below = rec[rec[:, engine.REC_MEASURED] != 1.0]
    assert len(below) >= 2 and not below[:, :12].any()
    assert (below[:, engine.REC_AREA] < engine.default_min_area(H, W)).all()
    ar = iset.area.cpu().numpy()
    assert np.array_equal(ar, np.array([int(mk.sum()) for mk in masks]))


def test_single_pass_overflow_falls_back(cuda_device):
    """Noise masks exceed the slab capacities (contours per instance, vertices): the exact two-pass path must take over."""
    import cv2
    rng = np.random.default_rng(9)
    masks = [(rng.random((96, 96)) < 0.45).astype(np.uint8) for _ in range(6)] + [np.zeros((96, 96), np.uint8)]
    iset = engine.from_masks(torch.as_tensor(np.stack(masks), device=cuda_device))
    engine.measure(iset, um_pix=1.0)
    assert iset.cstart_stride == 0      # fell back
    got = engine.contours_to_host(iset)
    for i, mk in enumerate(masks):
        ref = cv2.findContours(mk * 255, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)[0]
        assert len(got[i]) == len(ref)
        for a, b in zip(got[i], ref):
            assert np.array_equal(a, b[:, 0, :])


def _mask_lists(seed, n, H, W, dup=0.5):
    probs, boxes, scores, classes = _heads(seed, n, H, W, duplicate_frac=dup, rmin=5, rmax=18, margin=20)
    masks = d2_paste.paste_masks_in_image(probs, boxes, (H, W))
    return [m for m in masks], [np.float32(s) for s in scores], [int(c) for c in classes]


@pytest.mark.parametrize("seed,thr", [(21, 0.4), (22, 0.7), (23, 0.1)])
def test_dedup_smart_kept_sets(cuda_device, k4_path, seed, thr):
    H, W = 256, 256
    groups_ml = []
    for g in range(4):
        ml, sl, cl = _mask_lists(seed * 10 + g, 50 + 7 * g, H, W)
        if g % 2:
            ml.append(np.zeros((H, W), bool)); sl.append(np.float32(0.5123)); cl.append(0)
            t = np.zeros((H, W), bool); t[100, 20:110] = True
            ml.append(t); sl.append(np.float32(0.777)); cl.append(1)
        groups_ml.append((ml, sl, cl))
    allm = np.stack([m for ml, _, _ in groups_ml for m in ml])
    alls = np.array([s for _, sl, _ in groups_ml for s in sl], np.float32)
    allc = np.array([c for _, _, cl in groups_ml for c in cl], np.int32)
    offs = np.concatenate([[0], np.cumsum([len(ml) for ml, _, _ in groups_ml])])
    tm, ts, tc = _dev(cuda_device, allm.astype(np.uint8), alls, allc)
    iset = engine.from_masks(tm, scores=ts, classes=tc)
    engine.measure(iset, um_pix=1.0)
    kept = engine.dedup_smart(iset, engine.groups_from_offsets(offs, cuda_device), iou_threshold=thr).to_lists()
    for g, (ml, sl, cl) in enumerate(groups_ml):
        _, _, _, ref = dedup.deduplicate_masks_smart(ml, sl, cl, iou_threshold=thr, return_indices=True)
        assert [k - offs[g] for k in kept[g]] == ref, f"group {g}"


def test_dedup_quirks_q1_q2(cuda_device, k4_path):
    import cv2
    H = W = 128
    def disc(x, y, r=8):
        m = np.zeros((H, W), np.uint8); cv2.circle(m, (x, y), r, 1, -1); return m
    cases = [
        ([disc(20, 90), disc(20, 90)], [0.9, 0.8]),            # Q1: lower-left half, never de-duplicated
        ([disc(90, 20), disc(90, 20)], [0.9, 0.8]),            # upper-right half, de-duplicated
        ([disc(64, 64)] * 3, [0.7, 0.8, 0.9]),                 # Q2: keeps [0.9, 0.8]
        ([disc(64, 64)] * 3, [0.9, 0.8, 0.7]),                 # Q2: keeps [0.9]
    ]
    for masks, scores in cases:
        cl = [0] * len(masks)
        _, _, _, ref = dedup.deduplicate_masks_smart([m.astype(bool) for m in masks], [np.float32(s) for s in scores], cl,
                                                     iou_threshold=0.4, return_indices=True)
        tm, ts, tc = _dev(cuda_device, np.stack(masks), np.array(scores, np.float32), np.array(cl, np.int32))
        iset = engine.from_masks(tm, scores=ts, classes=tc)
        engine.measure(iset)
        got = engine.dedup_smart(iset, engine.groups_from_offsets([0, len(masks)], cuda_device), 0.4).to_lists()[0]
        assert got == ref


def test_inorder_dedup(cuda_device, k4_path):
    H, W = 256, 256
    ml, sl, cl = _mask_lists(77, 70, H, W, dup=0.8)
    ml.insert(5, np.zeros((H, W), bool)); sl.insert(5, np.float32(0.3)); cl.insert(5, 0)
    _, _, _, ref = dedup.greedy_inorder_dedup(ml, sl, 0, 0.5)
    tm, ts, tc = _dev(cuda_device, np.stack(ml).astype(np.uint8), np.array(sl, np.float32), np.array(cl, np.int32))
    iset = engine.from_masks(tm, scores=ts, classes=tc)
    got = engine.dedup_inorder(iset, engine.groups_from_offsets([0, len(ml)], cuda_device), 0.5).to_lists()[0]
    assert got == ref


@pytest.mark.parametrize("seed", [31, 32, 33])
def test_spatial_constraints(cuda_device, k4_path, seed):
    H, W = 256, 256
    rules = syn.POLYHIPES_RULES
    lists = [_mask_lists(seed * 10 + g, 60, H, W, dup=0.6) for g in range(3)]
    # make containment do something: a few small class-1 masks inside class-0 masks
    for ml, sl, cl in lists:
        for k in range(0, len(ml), 7):
            cl[k] = 0
        for k in range(3, len(ml), 7):
            big = ml[k - 3]
            ys, xs = np.nonzero(big)
            if len(ys) > 30:
                small = np.zeros_like(big); cy, cx = int(ys.mean()), int(xs.mean())
                small[cy - 2:cy + 3, cx - 2:cx + 3] = True
                ml[k] = small; cl[k] = 1
    allm = np.stack([m for ml, _, _ in lists for m in ml]).astype(np.uint8)
    alls = np.array([s for _, sl, _ in lists for s in sl], np.float32)
    allc = np.array([c for _, _, cl in lists for c in cl], np.int32)
    offs = np.concatenate([[0], np.cumsum([len(ml) for ml, _, _ in lists])])
    tm, ts, tc = _dev(cuda_device, allm, alls, allc)
    iset = engine.from_masks(tm, scores=ts, classes=tc)
    groups = engine.groups_from_offsets(offs, cuda_device)
    got_o = engine.overlap_rules(iset, groups, rules['overlap_rules']).to_lists()
    got_c = engine.containment_rules(iset, groups, rules['containment_rules'], 0.95).to_lists()
    got_all = engine.apply_spatial_constraints(iset, groups, rules).to_lists()
    got_c2 = engine.containment_rules(iset, groups, {1: 0, 0: 5}, 0.5).to_lists()     # chained rule + absent parent class
    for g, (ml, sl, cl) in enumerate(lists):
        n = len(ml)
        rem = spatial.filter_by_overlap_rules(ml, sl, cl, rules['overlap_rules'])
        assert [k - offs[g] for k in got_o[g]] == [i for i in range(n) if i not in rem]
        rem = spatial.filter_by_containment_rules(ml, sl, cl, rules['containment_rules'], 0.95)
        assert [k - offs[g] for k in got_c[g]] == [i for i in range(n) if i not in rem]
        _, _, _, idx = spatial.apply_spatial_constraints(ml, sl, cl, rules)
        assert [k - offs[g] for k in got_all[g]] == idx
        rem = spatial.filter_by_containment_rules(ml, sl, cl, {1: 0, 0: 5}, 0.5)
        assert [k - offs[g] for k in got_c2[g]] == [i for i in range(n) if i not in rem]


def test_fused_tiles_pipeline(cuda_device, k4_path):
    """BASELINE config 2/5 path on 3 tiles: paste -> measure -> de-dup 0.7 -> spatial constraints -> rows."""
    H, W = 384, 384
    tiles = [_heads(5000 + t, 70 + 5 * t, H, W, duplicate_frac=0.3, rmin=6, rmax=22, margin=25) for t in range(3)]
    probs = np.concatenate([t[0] for t in tiles]); boxes = np.concatenate([t[1] for t in tiles])
    scores = np.concatenate([t[2] for t in tiles]); classes = np.concatenate([t[3] for t in tiles])
    offs = np.concatenate([[0], np.cumsum([len(t[0]) for t in tiles])])
    tp, tb, ts, tc = _dev(cuda_device, probs, boxes, scores, classes)
    iset, kept, meas = engine.run_tiles(tp, tb, ts, tc, offs, H, W, um_pix=0.5, rules=syn.POLYHIPES_RULES, dedup_iou=0.7, frames=True)
    torch.cuda.synchronize()
    kept_l = kept.to_lists()
    host_rows = meas.rows_to_host()
    for t, (p, b, s, c) in enumerate(tiles):
        final, rows, _ = pipeline.run_tile(p, b, s, c, H, W, um_pix=0.5, rules=syn.POLYHIPES_RULES, dedup_iou=0.7)
        assert [k - offs[t] for k in kept_l[t]] == final, f"tile {t}"
        assert [k for k, _ in host_rows[t]] == kept_l[t]
        got = [r[:12] for _, rr in host_rows[t] for r in rr if r[engine.REC_MEASURED] == 1.0]
        assert len(got) == len(rows)
        for g, r in zip(got, rows):
            np.testing.assert_allclose(g, np.array([float(v) for v in r[3:15]]), rtol=REL_TOL, atol=0)


@pytest.mark.parametrize("host_inputs", [False, True])
def test_tile_pipeline_matches_single_batch_path(cuda_device, host_inputs):
    """The three-stream batched pipeline (bench.py's step) gives exactly the results of the one-batch path and the oracle."""
    H, W = 256, 256
    tiles = [_heads(7000 + t, 30 + 3 * t, H, W, duplicate_frac=0.3, rmin=5, rmax=18, margin=20) for t in range(7)]
    probs = np.concatenate([t[0] for t in tiles]); boxes = np.concatenate([t[1] for t in tiles])
    scores = np.concatenate([t[2] for t in tiles]); classes = np.concatenate([t[3] for t in tiles])
    offs = np.concatenate([[0], np.cumsum([len(t[0]) for t in tiles])])
    dev_in = _dev(cuda_device, probs, boxes, scores, classes)
    iset, kept, meas = engine.run_tiles(*dev_in, offs, H, W, um_pix=0.5, rules=syn.POLYHIPES_RULES, dedup_iou=0.7)
    ref_kept = kept.to_lists(); ref_rows = meas.rows_to_host()
    pipe = engine.TilePipeline(H, W, um_pix=0.5, rules=syn.POLYHIPES_RULES, dedup_iou=0.7, batches=3, device=cuda_device)
    if host_inputs:
        inputs = [torch.as_tensor(np.ascontiguousarray(a)).pin_memory() for a in (probs, boxes, scores, classes)]
    else:
        inputs = dev_in
    for _ in range(2):      # twice: buffers of the first run are recycled by the caching allocator
        res = pipe.run(*inputs, offs, to_host=host_inputs)
        torch.cuda.synchronize()
        seen = 0
        for r in res:
            t0, t1 = r["tiles"]
            kl = r["kept"].to_lists(); rows = r["meas"].rows_to_host()
            for g in range(t1 - t0):
                assert [k + r["inst0"] for k in kl[g]] == ref_kept[t0 + g]
                assert len(rows[g]) == len(ref_rows[t0 + g])
                for (ka, ra), (kb, rb) in zip(rows[g], ref_rows[t0 + g]):
                    assert ka + r["inst0"] == kb and np.array_equal(ra, rb)
                seen += 1
            if host_inputs:
                assert np.array_equal(r["host"]["records"].numpy(), r["meas"].records.cpu().numpy())
                assert np.array_equal(r["host"]["kept_idx"].numpy(), r["kept"].idx.cpu().numpy())
        assert seen == len(tiles)
    # and against the oracle for one tile
    p, b, s, c = tiles[4]
    final, rows, _ = pipeline.run_tile(p, b, s, c, H, W, um_pix=0.5, rules=syn.POLYHIPES_RULES, dedup_iou=0.7)
    assert [k - offs[4] for k in ref_kept[4]] == final


def test_tile_pipeline_sync_free_guards(cuda_device):
    """Runs after the first of a shard shape are enqueued without reading sizes back; when the new inputs need more room than the
    previous run left, the device-side guard trips, nothing is written, aborted() reports it and the re-run is exact again."""
    H, W = 256, 256
    def shard(seed, rmax):
        tiles = [_heads(seed + t, 40, H, W, duplicate_frac=0.3, rmin=5, rmax=rmax, margin=30) for t in range(4)]
        cat = [np.concatenate([t[k] for t in tiles]) for k in range(4)]
        offs = np.concatenate([[0], np.cumsum([len(t[0]) for t in tiles])])
        return _dev(cuda_device, *cat), offs
    pipe = engine.TilePipeline(H, W, um_pix=0.5, rules=syn.POLYHIPES_RULES, dedup_iou=0.7, batches=2, device=cuda_device)
    small, offs = shard(100, 9)
    big, offs_b = shard(200, 28)
    assert np.array_equal(offs, offs_b)                      # same shard shape, much larger particles
    def result(res):
        out = []
        for r in res:
            r["meas"].finalize()
            out.append((r["kept"].to_lists(), r["meas"].records.cpu().numpy().copy()))
        return out
    ref_small = result(pipe.run(*small, offs)); assert not pipe.aborted()          # exact-size path, records the hints
    again = result(pipe.run(*small, offs)); assert not pipe.aborted()               # sync-free path
    assert all(a[0] == b[0] and np.array_equal(a[1], b[1]) for a, b in zip(ref_small, again))
    pipe.run(*big, offs)                                                            # does not fit the hints
    assert pipe.aborted()
    got_big = result(pipe.run(*big, offs)); assert not pipe.aborted()               # exact-size path again
    fresh = engine.TilePipeline(H, W, um_pix=0.5, rules=syn.POLYHIPES_RULES, dedup_iou=0.7, batches=2, device=cuda_device, sync_free=False)
    ref_big = result(fresh.run(*big, offs))
    assert all(a[0] == b[0] and np.array_equal(a[1], b[1]) for a, b in zip(ref_big, got_big))
