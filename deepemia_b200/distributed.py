"""Multi-GPU pieces of the hot path (one process per GPU, torch.distributed; NCCL on GPUs, gloo in the CPU tests).

Most of the path shards with NO data-path collective (independent images / tiles: bench.py).  The exception is a
single large micrograph split over the GPUs (BASELINE config 3): every rank back-projects the tiles of its contiguous band,
but the reference's global deduplicate_masks_smart (src/functions/inference.py:2472) is a greedy pass whose outcome depends on
the score order and list order of ALL instances (Q2) — including the full-image pass, which overlaps every band.  The instances
are tiny once bit-packed (a 20 000-instance micrograph is ~10 MB), so the exchange is ONE variable-length all-gather of the
packed instance records, after which every rank replays the identical global greedy pass and measures its share of survivors.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import engine


def band_of_rank(n_tiles, rank, world):
    """Contiguous tile range [t0, t1) of `rank` (rank-major concatenation == global tile order)."""
    base, extra = divmod(n_tiles, world)
    t0 = rank * base + min(rank, extra)
    return t0, t0 + base + (1 if rank < extra else 0)


def _pad_to(t, n):
    if t.shape[0] == n:
        return t.contiguous()
    pad = torch.zeros((n - t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    return torch.cat([t, pad]).contiguous()


def all_gather_instances(iset, group=None):
    """All ranks contribute an InstanceSet of the same frame (scores and classes attached); every rank gets the rank-major
    concatenation.  Variable length: one all-gather of (n, crop words), then padded all-gathers of the record arrays."""
    world = dist.get_world_size(group)
    dev = iset.device
    n, words = iset.n, iset.total_crop_words
    sizes = torch.tensor([n, words], dtype=torch.int64, device=dev)
    all_sizes = [torch.empty_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    all_sizes = torch.stack(all_sizes).cpu().numpy()
    n_max, w_max = int(all_sizes[:, 0].max()), int(all_sizes[:, 1].max())
    scores = iset.scores if iset.scores is not None else torch.zeros(n, dtype=torch.float32, device=dev)
    classes = iset.classes if iset.classes is not None else torch.zeros(n, dtype=torch.int32, device=dev)
    payload = {
        "meta": iset.meta[:n], "crop_off": iset.crop_off[:n], "bbox": iset.bbox[:n], "area": iset.area[:n],
        "scores": scores[:n].to(torch.float32), "classes": classes[:n].to(torch.int32),
    }
    gathered = {}
    for k, t in payload.items():
        buf = _pad_to(t, max(n_max, 1))
        outs = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(outs, buf, group=group)
        gathered[k] = outs
    cbuf = _pad_to(iset.crops[:words], max(w_max, 1))
    couts = [torch.empty_like(cbuf) for _ in range(world)]
    dist.all_gather(couts, cbuf, group=group)
    parts = []
    for r in range(world):
        nr, wr = int(all_sizes[r, 0]), int(all_sizes[r, 1])
        if nr == 0:
            continue
        co = torch.cat([gathered["crop_off"][r][:nr], torch.tensor([wr], dtype=torch.int64, device=dev)])
        parts.append(engine.InstanceSet(n=nr, H=iset.H, W=iset.W, meta=gathered["meta"][r][:nr], crop_off=co, crops=couts[r][:max(wr, 1)],
                                        bbox=gathered["bbox"][r][:nr], area=gathered["area"][r][:nr], scores=gathered["scores"][r][:nr],
                                        classes=gathered["classes"][r][:nr], total_crop_words=wr))
    if not parts:
        return None, all_sizes[:, 0]
    return engine.concat(parts), all_sizes[:, 0]


def split_micrograph_pipeline(predictor, image, target_class, small_classes, confidence_threshold, tile_size=512, overlap_ratio=0.1,
                              upscale_factor=2.0, iou_threshold=0.7, edge_filter_enabled=True, class_specific_settings=None,
                              confidence_mode='auto', group=None):
    """tile_based_inference_pipeline (src/functions/inference.py:2299-2485) for ONE micrograph over all ranks of `group`:
    rank r runs the tiles of its band (rank 0 also the full-image pass), one all-gather of the packed instances, then the global
    deduplicate_masks_smart at 0.4 on every rank (identical replay).  Returns the same (masks, scores, classes) on every rank."""
    import cv2
    from .functions import inference as inf
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    h, w = image.shape[:2]
    tiles = inf.generate_tiles_with_overlap(image, tile_size, overlap_ratio)
    t0, t1 = band_of_rank(len(tiles), rank, world)
    parts = []
    if rank == 0:
        parts.append(inf._dev_run_class_specific_inference(predictor, image, target_class, small_classes, confidence_threshold,
                                                           iou_threshold=iou_threshold, class_specific_settings=class_specific_settings,
                                                           confidence_mode=confidence_mode))
    for tile_img, x_offset, y_offset in tiles[t0:t1]:
        tile_h, tile_w = tile_img.shape[:2]
        up = cv2.resize(tile_img, (int(tile_w * upscale_factor), int(tile_h * upscale_factor)), interpolation=cv2.INTER_LINEAR)
        t = inf._dev_run_class_specific_inference(predictor, up, target_class, small_classes, confidence_threshold,
                                                  iou_threshold=iou_threshold, class_specific_settings=class_specific_settings,
                                                  confidence_mode=confidence_mode)
        if not len(t):
            continue
        off = np.tile(np.array([[x_offset, y_offset]], np.int32), (len(t), 1))
        placed, edge = engine.resize_place(t.iset, tile_h, tile_w, h, w, off_xy=off, tile_size=tile_size, overlap_ratio=overlap_ratio)
        d = inf._Dev(placed, t.scores, t.classes, [bool] * len(t))
        if edge_filter_enabled:
            d = inf._select(d, np.nonzero(edge.cpu().numpy()[:len(t)] == 0)[0].tolist())
        parts.append(d)
    local = inf._concat(parts)
    dev = torch.device("cuda", torch.cuda.current_device())
    if len(local):
        inf._with_scores(local)
        mine = local.iset
        kinds = [0 if k is np.uint8 else 1 for k in local.dtypes]
    else:
        z = lambda *s, dt=torch.int32: torch.zeros(s, dtype=dt, device=dev)
        mine = engine.InstanceSet(n=0, H=h, W=w, meta=z(0, 8), crop_off=z(1, dt=torch.int64), crops=z(1), bbox=z(0, 4), area=z(0),
                                  scores=z(0, dt=torch.float32), classes=z(0), total_crop_words=0)
        kinds = []
    allset, counts = all_gather_instances(mine, group=group)
    all_kinds = [None] * world
    dist.all_gather_object(all_kinds, kinds, group=group)
    if allset is None:
        return [], [], []
    kinds = sum(all_kinds, [])
    keep = inf._dedup_smart_ids(allset, 0.4)
    sel = engine.select(allset, keep)
    scores = allset.scores.cpu().numpy()
    classes = allset.classes.cpu().numpy()
    d = inf._Dev(sel, [np.float32(scores[i]) for i in keep], [int(classes[i]) for i in keep], [np.uint8 if kinds[i] == 0 else bool for i in keep])
    return inf._lists(d)
