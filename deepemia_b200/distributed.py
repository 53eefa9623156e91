"""Multi-GPU pieces of the hot path (one process per GPU, torch.distributed; NCCL on GPUs, gloo in the CPU tests).

Most of the path shards with NO data-path collective (independent images / tiles: bench.py).  The exception is
a single large micrograph split over the GPUs (BASELINE config 3): every rank runs the batched per-tile flow on a contiguous band
of tiles (rank 0 also the full-image pass), but the reference's global deduplicate_masks_smart
(src/functions/inference.py:2472) is a greedy pass whose outcome depends on the score order AND the list order of ALL
instances (Q2) — including the full-image pass, which overlaps every band.  Bit-packed, the surviving instances of a whole
8192 x 8192 micrograph are a few MB, so the exchange is

    ONE all-gather of sizes  (instances, crop words, per-class list lengths of every rank)
    ONE all-gather of a byte-packed payload  (meta | bbox | area | score | class | crop words of the rank's surviving instances)

after which every rank holds the identical global instance set in the reference's list order (rank-major == full image, then the
tiles in generate_tiles_with_overlap order) and replays the identical global stages.  The pair stage of the sparse K4 path costs
~0.1 ms for a whole micrograph, so distributing it (two more collectives for a pair-list exchange) would not pay.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import engine

_FIELDS = (("meta", torch.int32, 8), ("bbox", torch.int32, 4), ("area", torch.int32, 1), ("scores", torch.float32, 1),
           ("classes", torch.int32, 1))


def band_of_rank(n_tiles, rank, world):
    """Contiguous tile range [t0, t1) of `rank` (rank-major concatenation == global tile order)."""
    base, extra = divmod(n_tiles, world)
    t0 = rank * base + min(rank, extra)
    return t0, t0 + base + (1 if rank < extra else 0)


def _bytes(t):
    t = t.contiguous().reshape(-1)
    return t.view(torch.uint8) if t.numel() else torch.zeros(0, dtype=torch.uint8, device=t.device)


def pack_instances(iset):
    """One uint8 buffer: the per-instance records followed by the crop words (the set must be compact: crop offsets 0 .. words)."""
    n, words = iset.n, int(iset.total_crop_words)
    dev = iset.device
    parts = []
    for name, dt, width in _FIELDS:
        t = getattr(iset, name)
        if t is None:
            t = torch.zeros((n, width) if width > 1 else (n,), dtype=dt, device=dev)
        parts.append(_bytes(t[:n].to(dt)))
    sizes = (iset.crop_off[1:n + 1] - iset.crop_off[:n]).to(torch.int32)        # 4-byte fields only: every view stays aligned
    parts.append(_bytes(sizes))
    parts.append(_bytes(iset.crops[:words]))
    return torch.cat(parts) if parts else torch.zeros(0, dtype=torch.uint8, device=dev)


def _payload_bytes(n, words):
    return n * (32 + 16 + 4 + 4 + 4 + 4) + words * 4


def all_gather_packed(iset, extra=(), group=None):
    """Every rank contributes a compact InstanceSet of the same frame; every rank gets the rank-major concatenation.
    Exactly two collectives: sizes (+ `extra` int64 values per rank, e.g. per-class list lengths), then one byte payload.
    Returns (global InstanceSet or None, sizes int64 [world, 2 + len(extra)])."""
    world = dist.get_world_size(group)
    dev = iset.device
    n, words = iset.n, int(iset.total_crop_words)
    mine = torch.tensor([n, words] + [int(v) for v in extra], dtype=torch.int64, device=dev)
    all_sizes = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(all_sizes, mine, group=group)                                   # collective 1
    all_sizes = torch.stack(all_sizes).cpu().numpy()
    nbytes = [_payload_bytes(int(r[0]), int(r[1])) for r in all_sizes]
    cap = max(max(nbytes), 16)
    buf = torch.zeros(cap, dtype=torch.uint8, device=dev)
    payload = pack_instances(iset)
    buf[:payload.numel()].copy_(payload)
    outs = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf, group=group)                                         # collective 2
    cols = {name: [] for name, _, _ in _FIELDS}
    offs, crops, base = [], [], 0
    for r in range(world):
        nr, wr = int(all_sizes[r, 0]), int(all_sizes[r, 1])
        if nr == 0:
            continue
        b = outs[r]
        pos = 0
        for name, dt, width in _FIELDS:
            nb = nr * width * 4
            t = b[pos:pos + nb].view(dt)
            cols[name].append(t.reshape(nr, width) if width > 1 else t)
            pos += nb
        sz = b[pos:pos + nr * 4].view(torch.int32).to(torch.int64)
        pos += nr * 4
        offs.append(torch.cumsum(sz, 0) - sz + base)
        crops.append(b[pos:pos + wr * 4].view(torch.int32))
        base += wr
    if not offs:
        return None, all_sizes
    n_all = int(all_sizes[:, 0].sum())
    crop_off = torch.cat(offs + [torch.tensor([base], dtype=torch.int64, device=dev)])
    g = engine.InstanceSet(n=n_all, H=iset.H, W=iset.W, meta=torch.cat(cols["meta"]).contiguous(), crop_off=crop_off.contiguous(),
                           crops=torch.cat(crops).contiguous() if base else torch.zeros(1, dtype=torch.int32, device=dev),
                           bbox=torch.cat(cols["bbox"]).contiguous(), area=torch.cat(cols["area"]).contiguous(),
                           scores=torch.cat(cols["scores"]).contiguous(), classes=torch.cat(cols["classes"]).contiguous(),
                           total_crop_words=base)
    return g, all_sizes


def split_micrograph(full_hb, tile_hb, tile_xy, image_hw, tile_size, overlap_ratio, params, edge_filter_enabled=True,
                     cross_class_iou=0.7, rules=None, um_pix=1.0, group=None, arena=None):
    """The per-image body of run_inference (tile_based_inference_pipeline per class :2299-2485, then :859-868 and the measurement
    loop) for ONE micrograph split over the ranks of `group`.

    full_hb : HeadBatch of the full-image pass (used on rank 0 only; None elsewhere is fine).
    tile_hb : HeadBatch of THIS rank's band of upscaled tiles (band_of_rank order), tile_xy: their (x, y) offsets.
    Returns dict(iset, per_class (lists after the 0.4 de-dup, one per class), kept (final list), meas) — identical on every rank."""
    from . import batched
    rank = dist.get_rank(group)
    dev = torch.device("cuda", torch.cuda.current_device())
    h, w = image_hw
    C = len(params)
    arena = arena or engine.Arena(dev)
    T = len(tile_xy)
    have_full = rank == 0 and full_hb is not None and full_hb.n > 0
    have_tiles = tile_hb is not None and tile_hb.U > 0 and tile_hb.n > 0
    ids_by_class = [[] for _ in range(C)]
    local = None
    if have_full or have_tiles:
        for _ in range(12):
            arena.begin()
            parts, kept_spaces = [], []
            if have_full:
                post_f, kept_f = batched.class_specific(full_hb, params, arena, tag="full")
                parts.append((post_f, kept_f, h, w, None, None))
            if have_tiles:
                post_t, kept_t = batched.class_specific(tile_hb, params, arena, tag="tile")
                xy = torch.as_tensor(np.ascontiguousarray(np.asarray(tile_xy, np.int32).reshape(T, 2)), device=dev)
                uoff = torch.as_tensor(np.asarray(tile_hb.unit_off, np.int32), device=dev)
                off_xy = engine.unit_broadcast(uoff, tile_hb.U, tile_hb.n, xy, 2)
                parts.append((post_t, kept_t, tile_size, tile_size, off_xy, (tile_size, overlap_ratio)))
            comb = engine.Combined([p[0].n for p in parts], h, w, dev)
            edge = {}
            for pi, (post, kept, th, tw, off_xy, e) in enumerate(parts):
                comb.plan(pi, post, th, tw, off_xy, alive=engine.mark_members(post, batched._whole(kept), -1))
                if e is not None and edge_filter_enabled:
                    edge[pi] = e
            local = comb.place(arena, tag="k3", edge=edge)
            lists = []
            for pi, (post, kept, th, tw, off_xy, e) in enumerate(parts):
                g = batched._whole(kept)
                if pi in edge:
                    g = engine.filter_flag(g, comb.edge[int(comb.start[pi]):], 0)
                lists.append(g)
            if arena.finish():
                break
        # surviving ids per class, in the reference's order (full-image members first, then the tiles in order)
        for pi, g in enumerate(lists):
            per_group = g.to_lists()                       # C sections x units
            U = len(per_group) // C
            for c in range(C):
                for u in range(U):
                    ids_by_class[c] += [int(comb.start[pi]) + i for i in per_group[c * U + u]]
    counts = [len(v) for v in ids_by_class]
    flat = [i for v in ids_by_class for i in v]
    if flat:
        mine = engine.select(local, flat)
    else:
        z = lambda *s, dt=torch.int32: torch.zeros(s, dtype=dt, device=dev)
        mine = engine.InstanceSet(n=0, H=h, W=w, meta=z(0, 8), crop_off=z(1, dt=torch.int64), crops=z(1), bbox=z(0, 4), area=z(0),
                                  scores=z(0, dt=torch.float32), classes=z(0), total_crop_words=0)
    allset, sizes = all_gather_packed(mine, extra=counts, group=group)
    if allset is None:
        return dict(iset=None, per_class=[[] for _ in range(C)], kept=[], meas=None)
    # global per-class lists: rank-major
    bases = np.concatenate([[0], np.cumsum(sizes[:, 0])])
    glists = []
    for c in range(C):
        ids = []
        for r in range(sizes.shape[0]):
            a = int(bases[r] + sizes[r, 2:2 + c].sum())
            ids += list(range(a, a + int(sizes[r, 2 + c])))
        glists.append(ids)
    engine.trace(allset)
    groups = engine.groups_from_lists(glists, dev)
    flag = allset.extra.get("overflow")
    sp = engine.GroupSpace([np.diff(groups.cap_off_host).astype(np.int64)], dev)
    engine.dedup_smart(allset, groups, 0.4, out=sp.section(0))
    if flag is not None and int(flag.item()):
        engine.trace(allset, single_pass=False)
        engine.dedup_smart(allset, groups, 0.4, out=sp.section(0))
    per_class = sp.section(0).to_lists()
    allc = engine.flatten(sp, [list(range(C))])
    kept = engine.dedup_smart(allset, allc, cross_class_iou)
    kept = engine.apply_spatial_constraints(allset, kept, rules)
    meas = engine.measure_list(allset, kept, um_pix=um_pix)
    return dict(iset=allset, per_class=per_class, kept=kept.to_lists()[0], meas=meas)
