"""GPU parity of K2 (bit-packed clean-up morphology) against the oracle's scipy restatement of
postprocess_masks (src/utils/mask_utils.py:38-84), process_masks_parallel (src/functions/inference.py:170-213) and
postprocess_masks_universal (inference.py:1739-1813).  Bit-exact."""
import cv2
import numpy as np
import pytest
import torch

from deepemia_b200 import engine, synthetic as syn
from oracle import morphology

pytestmark = pytest.mark.gpu


def _mask_zoo(H, W, seed):
    rng = np.random.default_rng(seed)
    out = [m for m in syn.masks_from_polys(syn.particle_field(rng, 40, H, W, rmin=5, rmax=25, margin=20), H, W)]
    m = np.zeros((H, W), np.uint8); cv2.circle(m, (60, 60), 25, 1, 3); out.append(m)                       # ring: hole to fill
    m = np.zeros((H, W), np.uint8); cv2.circle(m, (0, 90), 20, 1, 2); out.append(m)                        # arc against the frame edge
    m = np.zeros((H, W), np.uint8); m[0:12, 0:40] = 1; m[3:8, 5:30] = 0; out.append(m)                     # hole near the corner
    m = np.zeros((H, W), np.uint8); m[H - 10:, W - 37:] = 1; m[H - 6:H - 3, W - 20:W - 5] = 0; out.append(m)
    m = np.zeros((H, W), np.uint8); m[100, 20:120] = 1; out.append(m)                                      # 1-px line (erodes away)
    m = np.zeros((H, W), np.uint8); cv2.circle(m, (150, 150), 14, 1, -1); cv2.circle(m, (185, 150), 10, 1, -1); out.append(m)
    m = np.zeros((H, W), np.uint8); cv2.circle(m, (150, 60), 14, 1, -1); m[60, 164:170] = 1; cv2.circle(m, (182, 60), 12, 1, -1); out.append(m)
    m = (rng.random((H, W)) < 0.5).astype(np.uint8); m[:, W // 2:] = 0; m[H // 3:, :] = 0; out.append(m)   # noise: many holes / comps
    m = np.zeros((H, W), np.uint8); m[30:50, 31:34] = 1; m[30:33, 10:80] = 1; m[47:50, 10:80] = 1; m[30:50, 60:66] = 1; out.append(m)
    m = np.zeros((H, W), np.uint8); out.append(m)                                                           # empty
    m = np.ones((H, W), np.uint8); m[5:9, 5:9] = 0; out.append(m)                                            # whole frame with a hole
    # spiral: the flood needs many direction changes
    m = np.zeros((H, W), np.uint8)
    for k in range(6):
        cv2.rectangle(m, (100 + 6 * k, 170 + 6 * k), (200 - 6 * k, 250 - 6 * k), 1, 1)
        m[180 + 6 * k, 100 + 6 * k] = 0
    out.append(m)
    return out


def _bits(iset):
    return engine.unpack_masks(iset).cpu().numpy()


@pytest.mark.parametrize("shape", [(256, 256), (270, 301)])
def test_morph_chains_match_scipy(cuda_device, shape):
    H, W = shape
    masks = _mask_zoo(H, W, 7)
    iset = engine.from_masks(torch.as_tensor(np.stack(masks), device=cuda_device))
    # process_masks_parallel: fill -> erode -> dilate
    ref = morphology.process_masks_parallel(masks)
    out = engine.process_masks_parallel(iset)
    got = _bits(out)
    for i, r in enumerate(ref):
        assert np.array_equal(got[i], r), f"opening differs for mask {i}"
        assert int(out.area[i]) == int(r.sum())
    # single operators
    from scipy import ndimage as ndi
    for ops, fn in (([engine.MORPH_FILL], lambda m: ndi.binary_fill_holes(m).astype(np.uint8)),
                    ([engine.MORPH_ERODE], morphology.erosion),
                    ([engine.MORPH_FILL, engine.MORPH_DILATE, engine.MORPH_ERODE],
                     lambda m: morphology.erosion(morphology.dilation(ndi.binary_fill_holes(m).astype(np.uint8))))):
        got = _bits(engine.morph(iset, ops))
        for i, m in enumerate(masks):
            assert np.array_equal(got[i], fn(m)), f"ops {ops} differ for mask {i}"


@pytest.mark.parametrize("small", [True, False])
def test_postprocess_masks_universal(cuda_device, small):
    H, W = 256, 256
    masks = _mask_zoo(H, W, 8)
    iset = engine.from_masks(torch.as_tensor(np.stack(masks), device=cuda_device))
    groups = engine.groups_from_offsets([0, 20, len(masks)], cuda_device)
    out, kept = engine.postprocess_masks_universal(iset, groups, small)
    kl = kept.to_lists()
    got = _bits(out)
    for g, (a, b) in enumerate(((0, 20), (20, len(masks)))):
        area = H * W
        mcs = max(3, int(area * 0.000005)) if small else max(25, int(area * 0.0001))
        ref = []
        ref_idx = []
        from scipy import ndimage as ndi
        for i in range(a, b):
            f = ndi.binary_fill_holes(masks[i]).astype(np.uint8)
            fin = morphology.erosion(f) if small else morphology.dilation(morphology.erosion(f))
            if fin.sum() >= mcs:
                ref.append(fin.astype(bool)); ref_idx.append(i)
        oracle_list = morphology.postprocess_masks_universal([m.astype(bool) for m in masks[a:b]], (H, W), small)
        assert len(oracle_list) == len(ref)
        assert kl[g] == ref_idx
        for i, r in zip(ref_idx, oracle_list):
            assert np.array_equal(got[i].astype(bool), r)


@pytest.mark.parametrize("seed", [1, 2])
def test_postprocess_masks_overlap_and_components(cuda_device, seed):
    H, W = 256, 256
    rng = np.random.default_rng(seed)
    lists = []
    for g in range(3):
        polys = syn.particle_field(rng, 30, H, W, rmin=8, rmax=30, margin=30)
        # overlapping re-detections so that the first-come rule and the component test have work
        polys += [p + rng.uniform(-9, 9, 2) for p in polys[:15]]
        ms = syn.masks_from_polys(polys, H, W)
        m = np.zeros((H, W), np.uint8); cv2.circle(m, (128, 128), 40, 1, 4); ms.insert(3, m)
        m = np.zeros((H, W), np.uint8); m[120:136, 60:200] = 1; ms.append(m)           # bar cut in two by earlier masks
        lists.append(ms)
    allm = np.stack([m for ms in lists for m in ms])
    offs = np.concatenate([[0], np.cumsum([len(ms) for ms in lists])])
    iset = engine.from_masks(torch.as_tensor(allm, device=cuda_device))
    out, gated = engine.postprocess_masks(iset, engine.groups_from_offsets(offs, cuda_device))
    got = _bits(out)
    gl = gated.to_lists()
    for g, ms in enumerate(lists):
        ref = morphology.postprocess_masks(np.stack(ms), np.ones(len(ms), np.float32), (H, W))
        assert len(gl[g]) == len(ref)
        zeroed = 0
        for k, r in enumerate(ref):
            assert np.array_equal(got[offs[g] + k], r), f"group {g} member {k}"
            zeroed += int(r.sum() == 0)
        assert zeroed > 0


def test_postprocess_masks_column_gate(cuda_device):
    """Q5: fewer qualifying frame columns than masks truncates the list."""
    H, W = 64, 64
    ms = []
    for k in range(12):
        m = np.zeros((H, W), np.uint8); m[5 + 4 * k: 8 + 4 * k, 10:16] = 1; ms.append(m)    # only 6 columns ever set
    ref = morphology.postprocess_masks(np.stack(ms), np.ones(len(ms), np.float32), (H, W), min_crys_size=2)
    iset = engine.from_masks(torch.as_tensor(np.stack(ms), device=cuda_device))
    out, gated = engine.postprocess_masks(iset, engine.groups_from_offsets([0, len(ms)], cuda_device), min_crys_size=2)
    gl = gated.to_lists()[0]
    assert len(gl) == len(ref) == 6
    got = _bits(out)
    for k, r in enumerate(ref):
        assert np.array_equal(got[gl[k]], r)
    # threshold nobody reaches -> empty list
    ref = morphology.postprocess_masks(np.stack(ms), np.ones(len(ms), np.float32), (H, W), min_crys_size=1000)
    _, gated = engine.postprocess_masks(iset, engine.groups_from_offsets([0, len(ms)], cuda_device), min_crys_size=1000)
    assert ref == [] and gated.to_lists()[0] == []


def test_closing_grows_next_to_the_frame_border(cuda_device):
    """Out-of-frame neighbours are ignored by the erosion, so the closing of a mask that ends ONE pixel short of the frame border
    reaches the border (regression: the result does not fit the input's bbox crop)."""
    from scipy import ndimage as ndi
    H, W = 64, 96
    ms = []
    for (y0, y1, x0, x1) in ((1, 20, 30, 50), (40, H - 1, 30, 50), (20, 40, 1, 20), (20, 40, 70, W - 1), (1, H - 1, 1, W - 1),
                             (1, 10, 31, 33), (50, H - 1, 63, 65)):
        m = np.zeros((H, W), np.uint8); cv2.ellipse(m, ((x0 + x1) // 2, (y0 + y1) // 2), ((x1 - x0) // 2, (y1 - y0) // 2), 0, 0, 360, 1, -1)
        m[:y0] = 0; m[y1:] = 0; m[:, :x0] = 0; m[:, x1:] = 0
        ms.append(m)
    iset = engine.from_masks(torch.as_tensor(np.stack(ms), device=cuda_device))
    got = _bits(engine.morph(iset, [engine.MORPH_FILL, engine.MORPH_DILATE, engine.MORPH_ERODE]))
    grew = 0
    for i, m in enumerate(ms):
        ref = morphology.erosion(morphology.dilation(ndi.binary_fill_holes(m).astype(np.uint8)))
        assert np.array_equal(got[i], ref), f"mask {i}"
        grew += int(ref.sum() > 0 and (dedup_bbox(ref) != dedup_bbox(m)))
    assert grew > 0
    got = _bits(engine.morph(iset, [engine.MORPH_DILATE]))
    for i, m in enumerate(ms):
        assert np.array_equal(got[i], morphology.dilation(m)), f"dilation of mask {i}"


def dedup_bbox(m):
    ys, xs = np.nonzero(m)
    return (ys.min(), xs.min(), ys.max(), xs.max())
