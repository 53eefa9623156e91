"""Property tests (hypothesis) of the device path through the C ABI: the size-independent invariants SURVEY.md §4 asks for —
bit-packing round trips on arbitrary shapes, the INTER_NEAREST index maps for arbitrary size ratios, RLE decode(encode) = identity,
idempotence of the morphological operators and of the greedy de-duplication, linearity of the moment sums."""
import cv2
import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings, strategies as st

from deepemia_b200 import engine

pytestmark = pytest.mark.gpu
SET = dict(max_examples=30, deadline=None, derandomize=True, database=None,
           suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])      # the same examples on every run


def _random_masks(seed, n, H, W, density):
    rng = np.random.default_rng(seed)
    m = (rng.random((n, H, W)) < density).astype(np.uint8)
    for i in range(n):            # a few blobs so that the masks are not pure noise
        x, y, r = int(rng.integers(0, W)), int(rng.integers(0, H)), int(rng.integers(1, max(2, min(H, W) // 3)))
        cv2.circle(m[i], (x, y), r, 1, -1)
    if n > 2:
        m[0] = 0
        m[1] = 1
    return m


@settings(**SET)
@given(seed=st.integers(0, 10**6), n=st.integers(1, 9), H=st.integers(1, 150), W=st.integers(1, 300), density=st.sampled_from([0.0, 0.02, 0.5]))
def test_pack_unpack_round_trip(cuda_device, seed, n, H, W, density):
    """from_masks -> unpack_masks is the identity for any frame shape (widths that are not word multiples, single rows / columns);
    area = popcount, bbox = the extent of the set pixels."""
    m = _random_masks(seed, n, H, W, density)
    iset = engine.from_masks(torch.as_tensor(m, device=cuda_device))
    assert np.array_equal(engine.unpack_masks(iset).cpu().numpy(), m)
    area, bbox = iset.area.cpu().numpy(), iset.bbox.cpu().numpy()
    for i in range(n):
        assert area[i] == int(m[i].sum())
        ys, xs = np.nonzero(m[i])
        want = (ys.min(), xs.min(), ys.max(), xs.max()) if len(ys) else (-1, -1, -1, -1)
        assert tuple(bbox[i]) == want


@settings(**SET)
@given(seed=st.integers(0, 10**6), hs=st.integers(1, 140), ws=st.integers(1, 200), hd=st.integers(1, 140), wd=st.integers(1, 200))
def test_nearest_resize_any_ratio(cuda_device, seed, hs, ws, hd, wd):
    """K3's index maps equal cv2.resize(INTER_NEAREST) for arbitrary source / destination sizes (up, down, anisotropic, 1-pixel)."""
    m = _random_masks(seed, 3, hs, ws, 0.3)
    iset = engine.from_masks(torch.as_tensor(m, device=cuda_device))
    out, _ = engine.resize_place(iset, hd, wd, hd, wd)
    got = engine.unpack_masks(out).cpu().numpy()
    for i in range(len(m)):
        assert np.array_equal(got[i], cv2.resize(m[i], (wd, hd), interpolation=cv2.INTER_NEAREST))


@settings(**SET)
@given(seed=st.integers(0, 10**6), n=st.integers(1, 6), H=st.integers(1, 90), W=st.integers(1, 140), density=st.sampled_from([0.0, 0.1, 0.6, 1.0]))
def test_rle_decodes_to_the_mask(cuda_device, seed, n, H, W, density):
    """Decoding K6's (start, length) pairs — column-major, 1-indexed, as mask_utils.rle_encoding writes them — gives the mask back;
    the run lengths add up to the area; starts are strictly increasing and runs do not touch (maximal runs)."""
    m = _random_masks(seed, n, H, W, density)
    iset = engine.from_masks(torch.as_tensor(m, device=cuda_device))
    run_off, runs = engine.rle_encode(iset)
    run_off, runs = run_off.cpu().numpy(), runs.cpu().numpy().reshape(-1, 2)
    for i in range(n):
        r = runs[run_off[i]:run_off[i + 1]]
        flat = np.zeros(H * W, np.uint8)
        for s, l in r:
            flat[s - 1:s - 1 + l] = 1
        assert np.array_equal(flat.reshape(W, H).T, m[i])
        assert int(r[:, 1].sum()) == int(m[i].sum())
        assert np.all(r[1:, 0] > r[:-1, 0] + r[:-1, 1])


@settings(**SET)
@given(seed=st.integers(0, 10**6), H=st.integers(8, 120), W=st.integers(8, 160), op=st.sampled_from(["fill", "open", "close"]))
def test_morphology_idempotent(cuda_device, seed, H, W, op):
    """binary_fill_holes, opening (erode -> dilate) and closing (dilate -> erode) with the reference's 3x3 cross are idempotent."""
    m = _random_masks(seed, 4, H, W, 0.08)
    iset = engine.from_masks(torch.as_tensor(m, device=cuda_device))
    ops = {"fill": [engine.MORPH_FILL], "open": [engine.MORPH_ERODE, engine.MORPH_DILATE], "close": [engine.MORPH_DILATE, engine.MORPH_ERODE]}[op]
    once = engine.morph(iset, ops)
    twice = engine.morph(once, ops)
    a, b = engine.unpack_masks(once).cpu().numpy(), engine.unpack_masks(twice).cpu().numpy()
    assert np.array_equal(a, b)
    if op == "fill":
        assert np.all(a >= m)
    if op == "open":
        assert np.all(a <= m)


@settings(**SET)
@given(seed=st.integers(0, 10**6), n=st.integers(2, 40), thr=st.sampled_from([0.1, 0.4, 0.7]))
def test_dedup_idempotent_and_independent(cuda_device, seed, n, thr):
    """The survivors of the in-order iou() de-dup are pairwise below the threshold, and de-duplicating them again keeps all."""
    H, W = 96, 128
    rng = np.random.default_rng(seed)
    m = np.zeros((n, H, W), np.uint8)
    for i in range(n):
        cv2.circle(m[i], (int(rng.integers(10, W - 10)), int(rng.integers(10, H - 10))), int(rng.integers(3, 14)), 1, -1)
    iset = engine.from_masks(torch.as_tensor(m, device=cuda_device))
    g = engine.groups_from_offsets([0, n], cuda_device)
    kept = engine.dedup_inorder(iset, g, thr)
    ids = kept.to_lists()[0]
    again = engine.dedup_inorder(iset, kept, thr).to_lists()[0]
    assert again == ids and ids[0] == 0
    for a in range(len(ids)):
        for b in range(a):
            inter = int((m[ids[a]] & m[ids[b]]).sum())
            union = int((m[ids[a]] | m[ids[b]]).sum())
            assert not (union > 0 and inter / union > thr)


@settings(**SET)
@given(seed=st.integers(0, 10**6), H=st.integers(4, 100), W=st.integers(4, 150))
def test_raw_moments_are_additive(cuda_device, seed, H, W):
    """Raw moments are sums over pixels: for two disjoint masks m00..m03 of the union equal the sums of the parts."""
    rng = np.random.default_rng(seed)
    a = (rng.random((H, W)) < 0.3).astype(np.uint8)
    b = ((rng.random((H, W)) < 0.3) & (a == 0)).astype(np.uint8)
    iset = engine.from_masks(torch.as_tensor(np.stack([a, b, a | b]), device=cuda_device))
    mo = engine.moments(iset).cpu().numpy()
    assert np.array_equal(mo[0, :10] + mo[1, :10], mo[2, :10])
