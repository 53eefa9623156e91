"""GPU parity of K3 (nearest back-projection + edge test + placement), K6 (RLE), raw moments and the score-sorted iou() de-dup
against OpenCV / the oracle's restatement of src/functions/inference.py:2399-2420, :2522-2549, :2044-2054, :1964-1978 and
src/utils/mask_utils.py:17-35.  Bit-exact."""
import cv2
import numpy as np
import pytest
import torch

from deepemia_b200 import engine, synthetic as syn
from oracle import dedup, morphology, tiles

pytestmark = pytest.mark.gpu


def _masks(seed, n, H, W, rmin=6, rmax=28, margin=2):
    rng = np.random.default_rng(seed)
    polys = syn.particle_field(rng, n, H, W, rmin=rmin, rmax=rmax, margin=margin)
    return syn.masks_from_polys(polys, H, W)


@pytest.mark.parametrize("up,tile,overlap", [(2.0, 96, 0.25), (1.5, 100, 0.1), (3.5, 64, 0.2)])
def test_tile_back_projection(cuda_device, up, tile, overlap):
    """Masks predicted on upscaled tiles -> frame: resize NEAREST, is_edge_mask, placement with clipping at the image border."""
    h, w = 230, 310                                  # not a multiple of the stride: the last tiles hang over the image
    Hs = Ws = int(tile * up)
    origins = tiles.tile_origins(h, w, tile, overlap)
    src, off = [], []
    for t, (x, y) in enumerate(origins[:14]):
        ms = _masks(100 + t, 5, Hs, Ws, rmin=5, rmax=25, margin=-4)      # some touch the tile border
        ms.append(np.zeros((Hs, Ws), np.uint8))
        m = np.zeros((Hs, Ws), np.uint8); m[Hs - 9:, Ws - 30:] = 1; ms.append(m)     # bottom-right corner: clipped on edge tiles
        m = np.zeros((Hs, Ws), np.uint8); m[Hs // 2, 3:Ws - 3] = 1; ms.append(m)      # 1-px line: may vanish when downscaling
        src += ms; off += [(x, y)] * len(ms)
    iset = engine.from_masks(torch.as_tensor(np.stack(src), device=cuda_device))
    out, edge = engine.resize_place(iset, tile, tile, h, w, off_xy=np.array(off, np.int32), tile_size=tile, overlap_ratio=overlap)
    got = engine.unpack_masks(out).cpu().numpy().astype(bool)
    edge_dev = edge
    edge = edge.cpu().numpy()
    n_edge = 0
    for i, (m, (x, y)) in enumerate(zip(src, off)):
        down = cv2.resize(m, (tile, tile), interpolation=cv2.INTER_NEAREST).astype(bool)
        assert bool(edge[i]) == tiles.is_edge_mask(down, tile, overlap), f"edge flag of instance {i}"
        ref = tiles.back_project(m, tile, tile, x, y, h, w, tile, overlap, edge_filter_enabled=False)
        assert np.array_equal(got[i], ref), f"placed mask of instance {i}"
        assert int(out.area[i]) == int(ref.sum())
        bb = dedup.get_mask_bbox(ref)
        assert tuple(out.bbox[i].tolist()) == (tuple(int(v) for v in bb) if bb is not None else (-1, -1, -1, -1))
        n_edge += int(edge[i])
    assert 0 < n_edge < len(src)
    # the edge filter as a list operation
    groups = engine.groups_from_offsets([0, len(src)], cuda_device)
    kept = engine.filter_flag(groups, edge_dev, keep_value=0).to_lists()[0]
    assert kept == [i for i in range(len(src)) if not edge[i]]


@pytest.mark.parametrize("scale", [0.5, 0.7, 1.5, 2.5])
def test_scale_back_projection(cuda_device, scale):
    """process_single_scale: masks found on the scaled image are resized back with INTER_NEAREST (inference.py:2044-2054)."""
    h, w = 201, 333
    hs, ws = int(h * scale), int(w * scale)
    src = _masks(int(scale * 10), 30, hs, ws, rmin=4, rmax=20, margin=0)
    iset = engine.from_masks(torch.as_tensor(np.stack(src), device=cuda_device))
    out, _ = engine.resize_place(iset, h, w, h, w)
    got = engine.unpack_masks(out).cpu().numpy()
    for i, m in enumerate(src):
        ref = cv2.resize(m, (w, h), interpolation=cv2.INTER_NEAREST)
        assert np.array_equal(got[i], ref), f"instance {i}"


def test_rle_and_moments(cuda_device):
    H, W = 97, 130
    ms = _masks(5, 25, H, W, rmin=3, rmax=18, margin=0)
    m = np.zeros((H, W), np.uint8); m[:, 40:43] = 1; ms.append(m)              # full columns: runs continue across columns
    m = np.zeros((H, W), np.uint8); m[H - 1, 10] = 1; m[0, 11] = 1; ms.append(m)
    m = np.zeros((H, W), np.uint8); m[0, 0] = 1; m[H - 1, W - 1] = 1; ms.append(m)
    ms.append(np.zeros((H, W), np.uint8))
    ms.append(np.ones((H, W), np.uint8))
    iset = engine.from_masks(torch.as_tensor(np.stack(ms), device=cuda_device))
    run_off, runs = engine.rle_encode(iset)
    run_off = run_off.cpu().numpy(); runs = runs.cpu().numpy()
    mom = engine.moments01(iset).cpu().numpy()
    for i, m in enumerate(ms):
        ref = morphology.rle_encoding(m)
        got = runs[run_off[i]:run_off[i + 1]].reshape(-1).tolist()
        assert got == [int(v) for v in ref], f"RLE of instance {i}"
        mo = cv2.moments(m)
        assert (mom[i, 0], mom[i, 1], mom[i, 2]) == (int(mo["m00"]), int(mo["m10"]), int(mo["m01"]))


@pytest.mark.parametrize("k4", ["fused", "staged"])
def test_sorted_iou_dedup(cuda_device, k4):
    old = engine.FUSED_K4
    engine.FUSED_K4 = k4 == "fused"
    try:
        H, W = 200, 200
        rng = np.random.default_rng(3)
        polys = syn.particle_field(rng, 40, H, W, rmin=6, rmax=20, margin=15)
        polys += [p + rng.uniform(-2, 2, 2) for p in polys[:25]]
        ms = [m.astype(bool) for m in syn.masks_from_polys(polys, H, W)]
        ms.insert(7, np.zeros((H, W), bool))
        scores = syn.distinct_scores(rng, len(ms))
        # reference loop (inference.py:1964-1978)
        ref = []
        for idx in np.argsort(scores)[::-1]:
            if not any(dedup.iou(ms[idx], ms[j]) > 0.4 for j in ref):
                ref.append(int(idx))
        iset = engine.from_masks(torch.as_tensor(np.stack(ms).astype(np.uint8), device=cuda_device),
                                 scores=torch.as_tensor(scores, device=cuda_device))
        got = engine.dedup_sorted(iset, engine.groups_from_offsets([0, len(ms)], cuda_device), 0.4).to_lists()[0]
        assert got == ref
    finally:
        engine.FUSED_K4 = old


def test_select_and_concat(cuda_device):
    H, W = 128, 160
    a = _masks(1, 12, H, W); b = _masks(2, 9, H, W)
    ia = engine.from_masks(torch.as_tensor(np.stack(a), device=cuda_device))
    ib = engine.from_masks(torch.as_tensor(np.stack(b), device=cuda_device))
    cat = engine.concat([ia, ib])
    got = engine.unpack_masks(cat).cpu().numpy()
    assert np.array_equal(got, np.stack(a + b))
    pick = [20, 3, 3, 11, 0]
    sel = engine.select(cat, pick)
    assert np.array_equal(engine.unpack_masks(sel).cpu().numpy(), np.stack(a + b)[pick])
    assert torch.equal(sel.area, cat.area[torch.as_tensor(pick, device=cuda_device)])
