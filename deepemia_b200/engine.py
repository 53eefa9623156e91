"""Device-resident engine over libemia.so: instances stay bit-packed in HBM from the mask head to the CSV row.

PyTorch is used only for device memory, streams and (in bench.py) torch.distributed; every computation on the path is
a kernel of libemia.so called through the C ABI (include/emia.h) with raw tensor pointers.
"""
from dataclasses import dataclass, field
from typing import Optional

import numpy as np
import torch

from . import _lib

MASK_SIDE = 28
REC_FIELDS = 16
(REC_MAJOR, REC_MINOR, REC_ECC, REC_LENGTH, REC_WIDTH, REC_CED, REC_ASPECT, REC_CIRC, REC_CHORDS, REC_FERET, REC_ROUND,
 REC_SPHER, REC_AREA, REC_PERIM, REC_NVERT, REC_MEASURED) = range(16)

LAUNCHES = {"count": 0}   # kernels launched through the ABI (bench.py reports it as gpu_launches)
STAGE_TIMING = {"enabled": False, "events": []}   # (name, start, end) CUDA events when enabled (bench.py --breakdown)


class _stage:
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if STAGE_TIMING["enabled"]:
            self.a = torch.cuda.Event(enable_timing=True); self.b = torch.cuda.Event(enable_timing=True)
            self.a.record()
        return self

    def __exit__(self, *exc):
        if STAGE_TIMING["enabled"]:
            self.b.record()
            STAGE_TIMING["events"].append((self.name, self.a, self.b))
        return False


def stage_summary():
    torch.cuda.synchronize()
    out = {}
    for name, a, b in STAGE_TIMING["events"]:
        out.setdefault(name, []).append(a.elapsed_time(b))
    STAGE_TIMING["events"].clear()
    return {k: float(np.mean(v)) for k, v in out.items()}


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(dev):
    if not torch.cuda.is_available():
        raise _lib.EmiaError("deepemia_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device(dev if dev is not None else "cuda")


def pitch_words_for(W):
    """128-bit aligned bit rows: pitch = 16 * ceil(W / 128) bytes."""
    return 4 * ((W + 127) // 128)


def exclusive_scan_(t):
    """In-place exclusive scan of t[0:n] with the total in t[n] (int64, n+1 entries)."""
    lib = _lib.load()
    n = t.numel() - 1
    nb = int(lib.emia_scan_workspace_bytes(n))
    ws = torch.empty(nb, dtype=torch.uint8, device=t.device)
    _lib.check(lib.emia_exclusive_scan_i64(_ptr(t), n, _ptr(ws), nb, _stream()), "emia_exclusive_scan_i64")
    LAUNCHES["count"] += 3
    return t


@dataclass
class InstanceSet:
    """n bit-packed instances of one frame size (see "Instance layout" in include/emia.h)."""
    n: int
    H: int
    W: int
    meta: torch.Tensor            # int32 [n, 8]
    crop_off: torch.Tensor        # int64 [n+1]
    crops: torch.Tensor           # int32 words
    bbox: torch.Tensor            # int32 [n, 4]  (y_min, x_min, y_max, x_max) or -1
    area: torch.Tensor            # int32 [n]
    scores: Optional[torch.Tensor] = None   # float32 [n]
    classes: Optional[torch.Tensor] = None  # int32 [n]
    frames: Optional[torch.Tensor] = None   # int32 [slots, H, pitch_words]
    total_crop_words: int = 0
    # contours / morphometry (filled by measure())
    cont_off: Optional[torch.Tensor] = None
    pt_off: Optional[torch.Tensor] = None
    pts: Optional[torch.Tensor] = None
    cstart: Optional[torch.Tensor] = None
    records: Optional[torch.Tensor] = None
    rec_inst: Optional[torch.Tensor] = None
    perim0: Optional[torch.Tensor] = None
    n_records: int = 0
    cstart_stride: int = 0        # 0: packed cstart (cont_off[i] + i); else ints per instance slab
    extra: dict = field(default_factory=dict)

    @property
    def device(self):
        return self.meta.device

    @property
    def valid(self):
        return self.meta[:, 6] != 0


def paste(probs, boxes, H, W, scores=None, classes=None, scale_x=1.0, scale_y=1.0, frames=None, frame_slots=0,
          variant=0, crops_out=None):
    """K1.  probs [n,28,28] f32, boxes [n,4] f32 xyxy (mask-head outputs, device tensors) -> InstanceSet.
    frames: None (crops only), True (allocate n full frames) or a preallocated int32 [slots, H, pitch_words] ring."""
    lib = _lib.load()
    dev = probs.device
    n = int(probs.shape[0])
    assert probs.dtype == torch.float32 and boxes.dtype == torch.float32 and probs.is_contiguous() and boxes.is_contiguous()
    meta = torch.empty((n, 8), dtype=torch.int32, device=dev)
    crop_off = torch.empty(n + 1, dtype=torch.int64, device=dev)
    st = _stream()
    with _stage("paste_plan+scan"):
        _lib.check(lib.emia_paste_plan(_ptr(boxes), n, scale_x, scale_y, H, W, _ptr(meta), _ptr(crop_off), st), "emia_paste_plan")
        exclusive_scan_(crop_off)
    LAUNCHES["count"] += 1
    total = int(crop_off[n].item()) if n else 0
    if crops_out is not None:
        assert crops_out.numel() >= max(total, 1)
        crops = crops_out
    else:
        crops = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
    bbox = torch.empty((n, 4), dtype=torch.int32, device=dev)
    area = torch.empty(n, dtype=torch.int32, device=dev)
    pw = pitch_words_for(W)
    if frames is True:
        frames = torch.empty((max(n, 1), H, pw), dtype=torch.int32, device=dev)
    slots = int(frames.shape[0]) if frames is not None else 1
    if frames is not None:
        assert frames.shape[1] == H and frames.shape[2] == pw and frames.is_contiguous()
    with _stage("k1_paste"):
        _lib.check(lib.emia_paste_threshold_bitpack(_ptr(probs), _ptr(boxes), _ptr(meta), _ptr(crop_off), n, scale_x, scale_y, H, W,
                                                    _ptr(frames), slots, pw, _ptr(crops), _ptr(bbox), _ptr(area), variant, st),
                   "emia_paste_threshold_bitpack")
    LAUNCHES["count"] += 1
    return InstanceSet(n=n, H=H, W=W, meta=meta, crop_off=crop_off, crops=crops, bbox=bbox, area=area, scores=scores,
                       classes=classes, frames=frames, total_crop_words=total)


def from_masks(masks, scores=None, classes=None):
    """Import n byte masks [n,H,W] (uint8/bool device tensor, non-zero = set) as an InstanceSet (bbox crops)."""
    lib = _lib.load()
    dev = masks.device
    if masks.dtype == torch.bool:
        masks = masks.view(torch.uint8)
    assert masks.dtype == torch.uint8 and masks.is_contiguous() and masks.dim() == 3
    n, H, W = (int(s) for s in masks.shape)
    meta = torch.empty((n, 8), dtype=torch.int32, device=dev)
    crop_off = torch.empty(n + 1, dtype=torch.int64, device=dev)
    bbox = torch.empty((n, 4), dtype=torch.int32, device=dev)
    area = torch.empty(n, dtype=torch.int32, device=dev)
    st = _stream()
    _lib.check(lib.emia_mask_bbox(_ptr(masks), n, H, W, _ptr(meta), _ptr(crop_off), _ptr(bbox), _ptr(area), st), "emia_mask_bbox")
    exclusive_scan_(crop_off)
    total = int(crop_off[n].item()) if n else 0
    crops = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
    _lib.check(lib.emia_mask_pack(_ptr(masks), n, H, W, _ptr(meta), _ptr(crop_off), _ptr(crops), st), "emia_mask_pack")
    LAUNCHES["count"] += 2
    return InstanceSet(n=n, H=H, W=W, meta=meta, crop_off=crop_off, crops=crops, bbox=bbox, area=area, scores=scores,
                       classes=classes, total_crop_words=total)


def unpack_masks(iset, idx=None):
    """Full-frame byte masks [k,H,W] (0/1) of the selected instances (drop-in list-of-arrays surface)."""
    lib = _lib.load()
    dev = iset.device
    if idx is None:
        k = iset.n
        idx_t = None
    else:
        idx_t = torch.as_tensor(idx, dtype=torch.int32, device=dev).contiguous()
        k = int(idx_t.numel())
    out = torch.empty((k, iset.H, iset.W), dtype=torch.uint8, device=dev)
    _lib.check(lib.emia_mask_unpack(_ptr(iset.crops), _ptr(iset.meta), _ptr(iset.crop_off), _ptr(idx_t), k, iset.H, iset.W,
                                    _ptr(out), _stream()), "emia_mask_unpack")
    LAUNCHES["count"] += 1
    return out


MORPH_FILL, MORPH_ERODE, MORPH_DILATE = 1, 2, 3


def _pad_plan(iset):
    """Padded-plane offsets ((ch+2) x (cw+2) words per instance) and the 3-plane work buffer of the K2 kernels."""
    cached = iset.extra.get("pad_plan")
    if cached is not None:
        return cached
    lib = _lib.load()
    pad_off = torch.empty(iset.n + 1, dtype=torch.int64, device=iset.device)
    _lib.check(lib.emia_morph_plan(_ptr(iset.meta), iset.n, _ptr(pad_off), _stream()), "emia_morph_plan")
    exclusive_scan_(pad_off)
    LAUNCHES["count"] += 1
    total = int(pad_off[iset.n].item()) if iset.n else 0
    work = torch.empty(max(3 * total, 1), dtype=torch.int32, device=iset.device)
    iset.extra["pad_plan"] = (pad_off, work)
    return pad_off, work


def _derived(iset, crops):
    """A new InstanceSet with the geometry (meta, crop_off) of `iset` and new crop bits; bbox/area recomputed."""
    lib = _lib.load()
    bbox = torch.empty_like(iset.bbox)
    area = torch.empty_like(iset.area)
    _lib.check(lib.emia_crop_stats(_ptr(crops), _ptr(iset.meta), _ptr(iset.crop_off), iset.n, _ptr(bbox), _ptr(area), _stream()),
               "emia_crop_stats")
    LAUNCHES["count"] += 1
    out = InstanceSet(n=iset.n, H=iset.H, W=iset.W, meta=iset.meta, crop_off=iset.crop_off, crops=crops, bbox=bbox, area=area,
                      scores=iset.scores, classes=iset.classes, total_crop_words=iset.total_crop_words)
    if "pad_plan" in iset.extra:
        out.extra["pad_plan"] = iset.extra["pad_plan"]
    return out


def morph(iset, ops):
    """K2: apply `ops` (MORPH_FILL / MORPH_ERODE / MORPH_DILATE, 3x3 cross, at most 4) to every instance.  The result has
    the geometry of the input (every chain the reference uses — closing, opening, erosion — stays inside the crop)."""
    lib = _lib.load()
    ops = np.ascontiguousarray(ops, dtype=np.int32)
    pad_off, work = _pad_plan(iset)
    crops = torch.empty_like(iset.crops)
    if iset.n:
        with _stage("k2_morph"):
            _lib.check(lib.emia_morph(_ptr(iset.crops), _ptr(iset.meta), _ptr(iset.crop_off), iset.n, iset.H, iset.W, ops.ctypes.data,
                                      len(ops), _ptr(pad_off), _ptr(work), _ptr(crops), _stream()), "emia_morph")
        LAUNCHES["count"] += 1
    return _derived(iset, crops)


def overlap_first_come(iset, groups):
    """Tail of postprocess_masks (src/utils/mask_utils.py:77-82): list member k loses the pixels earlier members cover, then is
    zeroed if it has more than one 8-connected component.  Members keep their list position (Q6)."""
    lib = _lib.load()
    pad_off, work = _pad_plan(iset)
    crops = iset.crops.clone()
    if iset.n and groups.total_cap:
        with _stage("k2_overlap_first_come"):
            _lib.check(lib.emia_overlap_first_come(_ptr(iset.crops), _ptr(iset.meta), _ptr(iset.crop_off), _ptr(iset.bbox),
                                                   _ptr(groups.cap_off), groups.G, groups.total_cap, _ptr(groups.length),
                                                   _ptr(groups.idx), _ptr(pad_off), _ptr(work), _ptr(crops), _stream()),
                       "emia_overlap_first_come")
        LAUNCHES["count"] += 1
    return _derived(iset, crops)


def filter_area(iset, groups, min_area):
    lib = _lib.load()
    out = _new_groups_like(groups)
    _lib.check(lib.emia_group_filter_area(_ptr(groups.cap_off), groups.G, _ptr(groups.length), _ptr(groups.idx), _ptr(iset.area),
                                          int(min_area), _ptr(out.length), _ptr(out.idx), _stream()), "emia_group_filter_area")
    LAUNCHES["count"] += 1
    return out


def column_gate(iset, groups, min_size):
    lib = _lib.load()
    out = _new_groups_like(groups)
    _lib.check(lib.emia_column_gate(_ptr(iset.crops), _ptr(iset.meta), _ptr(iset.crop_off), _ptr(groups.cap_off), groups.G,
                                    _ptr(groups.length), _ptr(groups.idx), iset.W, int(min_size), _ptr(out.length), _ptr(out.idx),
                                    _stream()), "emia_column_gate")
    LAUNCHES["count"] += 1
    return out


def postprocess_masks(iset, groups, min_crys_size=2):
    """postprocess_masks (src/utils/mask_utils.py:38-84) on every group: column gate (Q5) -> fill holes -> closing ->
    first-come overlap removal -> multi-component masks zeroed (kept in the list, Q6).  Returns (InstanceSet, Groups)."""
    gated = column_gate(iset, groups, min_crys_size)
    closed = morph(iset, [MORPH_FILL, MORPH_DILATE, MORPH_ERODE])
    return overlap_first_come(closed, gated), gated


def process_masks_parallel(iset):
    """process_masks_parallel (src/functions/inference.py:170-213): fill holes -> erosion(disk 1) -> dilation(disk 1)."""
    return morph(iset, [MORPH_FILL, MORPH_ERODE, MORPH_DILATE])


def postprocess_masks_universal(iset, groups, is_small_class, min_crys_size=None):
    """postprocess_masks_universal (src/functions/inference.py:1739-1813).  Returns (InstanceSet, Groups of survivors)."""
    if min_crys_size is None:
        a = iset.H * iset.W
        min_crys_size = max(3, int(a * 0.000005)) if is_small_class else max(25, int(a * 0.0001))
    out = morph(iset, [MORPH_FILL, MORPH_ERODE] if is_small_class else [MORPH_FILL, MORPH_ERODE, MORPH_DILATE])
    return out, filter_area(out, groups, min_crys_size)


CAP_CONTOURS = 8   # contours per instance held by the single-pass slab layout


def measure(iset, um_pix=1.0, min_area=None, single_pass=True):
    """K5: external contours + morphometry records for every instance of the set (results stay on the device).
    single_pass: follow the borders once into bounded per-instance slabs; any overflow falls back to the exact
    count -> scan -> store path."""
    lib = _lib.load()
    dev = iset.device
    n = iset.n
    if min_area is None:
        min_area = max(5, iset.H * iset.W * 0.000005 * 0.05)     # src/functions/inference.py:1178-1184
    st = _stream()
    marks = torch.empty(max(2 * iset.total_crop_words, 1), dtype=torch.int32, device=dev)
    perim0 = torch.empty(max(n, 1), dtype=torch.float64, device=dev)
    done = False
    if single_pass and n:
        sizes = torch.empty((3, n + 1), dtype=torch.int64, device=dev)     # rows: n_contours, pt capacity, scratch bytes
        flag = torch.zeros(1, dtype=torch.int32, device=dev)
        with _stage("k5_plan+scan"):
            _lib.check(lib.emia_contour_trace_plan(_ptr(iset.meta), n, _ptr(sizes[1]), st), "emia_contour_trace_plan")
            exclusive_scan_(sizes[1])
        cap_total = iset.extra.get("pt_cap_total")
        if cap_total is None:
            # capacity is a pure function of the crop sizes: 4 * (sum ch + 32 * sum cw) + 32 * n_live; read it once
            cap_total = int(sizes[1, n].item())
        pts = torch.empty(max(cap_total, 1), dtype=torch.int32, device=dev)
        cstart = torch.empty(n * (CAP_CONTOURS + 1) + 1, dtype=torch.int32, device=dev)
        with _stage("k5_trace"):
            _lib.check(lib.emia_contour_trace_slab(_ptr(iset.crops), _ptr(iset.meta), _ptr(iset.crop_off), n, _ptr(marks), _ptr(sizes[1]),
                                                   CAP_CONTOURS, _ptr(pts), _ptr(cstart), _ptr(sizes[0]), _ptr(sizes[2]), _ptr(flag),
                                                   st), "emia_contour_trace_slab")
        with _stage("k5_scans"):
            exclusive_scan_(sizes[0])
            exclusive_scan_(sizes[2])
        LAUNCHES["count"] += 2
        tot = torch.stack([sizes[0, n], sizes[2, n], flag[0].to(torch.int64)]).tolist()
        n_rec, n_scr, overflow = int(tot[0]), int(tot[1]), int(tot[2])
        if overflow == 0:
            records = torch.empty((max(n_rec, 1), REC_FIELDS), dtype=torch.float64, device=dev)
            rec_inst = torch.empty(max(n_rec, 1), dtype=torch.int32, device=dev)
            scratch = torch.empty(max(n_scr, 16), dtype=torch.uint8, device=dev)
            with _stage("k5_measure"):
                _lib.check(lib.emia_contour_measure_stored(_ptr(iset.meta), n, _ptr(sizes[0]), _ptr(sizes[1]), _ptr(cstart),
                                                           CAP_CONTOURS + 1, _ptr(sizes[2]), float(um_pix), float(min_area), _ptr(pts),
                                                           _ptr(records), _ptr(rec_inst), _ptr(perim0), _ptr(scratch), st),
                           "emia_contour_measure_stored")
            LAUNCHES["count"] += 1
            iset.cstart_stride = CAP_CONTOURS + 1
            done = True
    if not done:
        sizes = torch.empty((3, n + 1), dtype=torch.int64, device=dev)     # rows: n_contours, n_points, scratch bytes
        with _stage("k5_count"):
            _lib.check(lib.emia_contour_count(_ptr(iset.crops), _ptr(iset.meta), _ptr(iset.crop_off), n, _ptr(marks), _ptr(sizes[0]),
                                              _ptr(sizes[1]), _ptr(sizes[2]), st), "emia_contour_count")
        with _stage("k5_scans"):
            for r in range(3):
                exclusive_scan_(sizes[r])
        LAUNCHES["count"] += 1
        totals = sizes[:, n].tolist() if n else [0, 0, 0]
        n_rec, n_pts, n_scr = int(totals[0]), int(totals[1]), int(totals[2])
        pts = torch.empty(max(n_pts, 1), dtype=torch.int32, device=dev)
        cstart = torch.empty(n_rec + n + 1, dtype=torch.int32, device=dev)
        records = torch.empty((max(n_rec, 1), REC_FIELDS), dtype=torch.float64, device=dev)
        rec_inst = torch.empty(max(n_rec, 1), dtype=torch.int32, device=dev)
        scratch = torch.empty(max(n_scr, 16), dtype=torch.uint8, device=dev)
        with _stage("k5_trace+measure"):
            _lib.check(lib.emia_contour_measure(_ptr(iset.crops), _ptr(iset.meta), _ptr(iset.crop_off), n, _ptr(marks), _ptr(sizes[0]),
                                                _ptr(sizes[1]), _ptr(sizes[2]), float(um_pix), float(min_area), _ptr(pts), _ptr(cstart),
                                                _ptr(records), _ptr(rec_inst), _ptr(perim0), _ptr(scratch), st), "emia_contour_measure")
        LAUNCHES["count"] += 2
        iset.cstart_stride = 0
    iset.cont_off, iset.pt_off = sizes[0], sizes[1]
    iset.pts, iset.cstart, iset.records, iset.rec_inst, iset.perim0 = pts, cstart, records[:n_rec], rec_inst[:n_rec], perim0
    iset.n_records = n_rec
    iset.extra["um_pix"] = um_pix
    iset.extra["min_area"] = min_area
    iset.extra.pop("n_contours", None)
    return iset


def contours_to_host(iset):
    """Host copy of the contour vertex lists: list (per instance) of lists (OpenCV order) of int32 [k,2] arrays."""
    pts = iset.pts.cpu().numpy().view(np.uint32)
    cont_off = iset.cont_off.cpu().numpy(); pt_off = iset.pt_off.cpu().numpy(); cstart = iset.cstart.cpu().numpy()
    out = []
    for i in range(iset.n):
        nc = int(cont_off[i + 1] - cont_off[i])
        base = i * iset.cstart_stride if iset.cstart_stride else cont_off[i] + i
        cs = cstart[base: base + nc + 1]
        cl = []
        for j in range(nc):
            k = nc - 1 - j
            p = pts[pt_off[i] + cs[k]: pt_off[i] + cs[k + 1]]
            cl.append(np.stack([p & 0xFFFF, p >> 16], 1).astype(np.int32))
        out.append(cl)
    return out


@dataclass
class Groups:
    """G lists of instance ids (one per image / tile / de-dup call)."""
    cap_off_host: np.ndarray      # int32 [G+1]
    cap_off: torch.Tensor         # int32 [G+1] (device)
    length: torch.Tensor          # int32 [G]
    idx: torch.Tensor             # int32 [cap_off[G]]

    @property
    def G(self):
        return len(self.cap_off_host) - 1

    @property
    def total_cap(self):
        return int(self.cap_off_host[-1])

    def to_lists(self):
        ln = self.length.cpu().numpy()
        ix = self.idx.cpu().numpy()
        return [ix[self.cap_off_host[g]: self.cap_off_host[g] + ln[g]].tolist() for g in range(self.G)]


def groups_from_offsets(offsets, device):
    """Identity lists: group g = instances [offsets[g], offsets[g+1])."""
    off = np.asarray(offsets, dtype=np.int32)
    cap_off = torch.as_tensor(off, device=device)
    length = torch.as_tensor(np.diff(off).astype(np.int32), device=device)
    idx = torch.arange(int(off[-1]), dtype=torch.int32, device=device) + int(off[0])
    off0 = (off - off[0]).astype(np.int32)
    return Groups(cap_off_host=off0, cap_off=torch.as_tensor(off0, device=device), length=length, idx=idx)


def groups_from_lists(lists, device):
    off = np.zeros(len(lists) + 1, np.int32)
    off[1:] = np.cumsum([len(l) for l in lists])
    flat = np.concatenate([np.asarray(l, np.int32) for l in lists]) if off[-1] else np.zeros(0, np.int32)
    return Groups(cap_off_host=off, cap_off=torch.as_tensor(off, device=device),
                  length=torch.as_tensor(np.diff(off).astype(np.int32), device=device),
                  idx=torch.as_tensor(flat, device=device))


_ws_cache = {}


def _workspace(groups, device):
    lib = _lib.load()
    key = (groups.cap_off_host.tobytes(), str(device))
    nbytes = _ws_cache.get(key)
    if nbytes is None:
        host = np.ascontiguousarray(groups.cap_off_host, dtype=np.int32)
        nbytes = int(lib.emia_group_workspace_bytes(host.ctypes.data, groups.G))
        _ws_cache.clear()
        _ws_cache[key] = nbytes
    return torch.empty(nbytes, dtype=torch.uint8, device=device), nbytes


def _new_groups_like(g):
    return Groups(cap_off_host=g.cap_off_host, cap_off=g.cap_off, length=torch.empty_like(g.length), idx=torch.empty_like(g.idx))


def dedup_smart(iset, groups, iou_threshold=0.4, max_aspect_ratio=None):
    """deduplicate_masks_smart (src/functions/inference.py:2552) on every group.  Needs measure() first (compactness)."""
    lib = _lib.load()
    assert iset.perim0 is not None, "run measure() before dedup_smart (the pre-filter needs contour perimeters)"
    ws, nb = _workspace(groups, iset.device)
    out = _new_groups_like(groups)
    ncont = iset.extra.get("n_contours")
    if ncont is None:
        ncont = (iset.cont_off[1:] - iset.cont_off[:-1]).contiguous()
        iset.extra["n_contours"] = ncont
    with _stage("k4_dedup_smart"):
      _lib.check(lib.emia_dedup_smart(_ptr(iset.crops), _ptr(iset.meta), _ptr(iset.crop_off), _ptr(iset.bbox), _ptr(iset.area),
                                    _ptr(iset.perim0), _ptr(ncont), _ptr(iset.scores), _ptr(iset.classes), _ptr(groups.cap_off),
                                    groups.G, groups.total_cap, _ptr(groups.length), _ptr(groups.idx), float(iou_threshold),
                                    float(max_aspect_ratio) if max_aspect_ratio else 0.0, _ptr(out.length), _ptr(out.idx),
                                    _ptr(ws), nb, _stream()), "emia_dedup_smart")
    LAUNCHES["count"] += 6
    return out


def dedup_inorder(iset, groups, iou_threshold):
    """Greedy in-order de-dup with iou() (src/functions/inference.py:1453-1459)."""
    lib = _lib.load()
    ws, nb = _workspace(groups, iset.device)
    out = _new_groups_like(groups)
    _lib.check(lib.emia_dedup_inorder(_ptr(iset.crops), _ptr(iset.meta), _ptr(iset.crop_off), _ptr(iset.bbox), _ptr(iset.area),
                                      _ptr(groups.cap_off), groups.G, groups.total_cap, _ptr(groups.length), _ptr(groups.idx),
                                      float(iou_threshold), _ptr(out.length), _ptr(out.idx), _ptr(ws), nb, _stream()),
               "emia_dedup_inorder")
    LAUNCHES["count"] += 6
    return out


def overlap_rules(iset, groups, rules):
    """filter_by_overlap_rules (src/utils/spatial_constraints.py:192).  rules: {class: {allow_overlap, max_iou_threshold}}."""
    lib = _lib.load()
    if not rules:
        return groups
    ncls = max(int(c) for c in rules) + 1
    active = np.zeros(ncls, np.int32)
    max_iou = np.zeros(ncls, np.float64)
    for c, r in rules.items():
        allow = r.get('allow_overlap', True)
        mi = r.get('max_iou_threshold', 0.5)
        if allow and mi >= 0.9:
            continue
        active[int(c)] = 1
        max_iou[int(c)] = mi
    dev = iset.device
    act_t, mi_t = torch.as_tensor(active, device=dev), torch.as_tensor(max_iou, device=dev)
    ws, nb = _workspace(groups, dev)
    out = _new_groups_like(groups)
    with _stage("k4_overlap_rules"):
      _lib.check(lib.emia_overlap_rules(_ptr(iset.crops), _ptr(iset.meta), _ptr(iset.crop_off), _ptr(iset.bbox), _ptr(iset.area),
                                      _ptr(iset.scores), _ptr(iset.classes), _ptr(groups.cap_off), groups.G, groups.total_cap,
                                      _ptr(groups.length), _ptr(groups.idx), _ptr(act_t), _ptr(mi_t), ncls, _ptr(out.length),
                                      _ptr(out.idx), _ptr(ws), nb, _stream()), "emia_overlap_rules")
    LAUNCHES["count"] += 6
    return out


def containment_rules(iset, groups, rules, threshold=0.95):
    """filter_by_containment_rules (src/utils/spatial_constraints.py:280).  rules: {child_class: parent_class} (ordered)."""
    lib = _lib.load()
    if not rules:
        return groups
    child = np.asarray([int(c) for c in rules.keys()], np.int32)
    parent = np.asarray([int(p) for p in rules.values()], np.int32)
    ws, nb = _workspace(groups, iset.device)
    out = _new_groups_like(groups)
    with _stage("k4_containment"):
      _lib.check(lib.emia_containment_rules(_ptr(iset.crops), _ptr(iset.meta), _ptr(iset.crop_off), _ptr(iset.bbox), _ptr(iset.area),
                                          _ptr(iset.classes), _ptr(groups.cap_off), groups.G, groups.total_cap, _ptr(groups.length),
                                          _ptr(groups.idx), child.ctypes.data, parent.ctypes.data, len(child), float(threshold),
                                          _ptr(out.length), _ptr(out.idx), _ptr(ws), nb, _stream()), "emia_containment_rules")
    LAUNCHES["count"] += 1 + 2 * len(child)
    return out


def apply_spatial_constraints(iset, groups, rules):
    """apply_spatial_constraints (src/utils/spatial_constraints.py:401) with an explicit rule dict."""
    if not rules or not rules.get('enabled', False):
        return groups
    g = overlap_rules(iset, groups, rules.get('overlap_rules', {}))
    return containment_rules(iset, g, rules.get('containment_rules', {}), rules.get('containment_threshold', 0.95))


def pair_counts(iset, pa, pb):
    """(intersection, area_a, area_b) int64 [k,3] for explicit instance pairs."""
    lib = _lib.load()
    dev = iset.device
    pa = torch.as_tensor(pa, dtype=torch.int32, device=dev).contiguous()
    pb = torch.as_tensor(pb, dtype=torch.int32, device=dev).contiguous()
    out = torch.empty((pa.numel(), 3), dtype=torch.int64, device=dev)
    _lib.check(lib.emia_pair_counts(_ptr(iset.crops), _ptr(iset.meta), _ptr(iset.crop_off), _ptr(iset.area), _ptr(pa), _ptr(pb),
                                    pa.numel(), _ptr(out), _stream()), "emia_pair_counts")
    LAUNCHES["count"] += 1
    return out


def run_tiles(probs, boxes, scores, classes, tile_offsets, H, W, um_pix=0.5, rules=None, dedup_iou=0.7, frames=None,
              variant=0, scale_x=1.0, scale_y=1.0):
    """The fused hot path over many tiles at once (BASELINE configs 2 and 5): paste -> contours/morphometry ->
    deduplicate_masks_smart -> spatial constraints.  Everything stays on the device.  Returns (InstanceSet, Groups)."""
    iset = paste(probs, boxes, H, W, scores=scores, classes=classes, scale_x=scale_x, scale_y=scale_y, frames=frames,
                 variant=variant)
    measure(iset, um_pix=um_pix)
    groups = groups_from_offsets(tile_offsets, iset.device)
    kept = dedup_smart(iset, groups, iou_threshold=dedup_iou)
    kept = apply_spatial_constraints(iset, kept, rules)
    return iset, kept
