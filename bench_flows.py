"""bench_flows.py — BASELINE configs 2, 3a, 3b and 4 as bench workloads (imported by bench.py).

Each workload = synthetic head outputs -> one batched, sync-free flow of deepemia_b200.batched per step.
    value        : heads entering K1 per second, inputs resident in HBM, CUDA events over K steps, no host synchronisation
                   between steps (validity of every step is checked afterwards from its abort flag)
    e2e          : the same flow from pinned HOST buffers: H2D of the step's head outputs, the flow, D2H of the kept lists and the
                   measurement records, host synchronisation — every step
    cpu_baseline : oracle/flows.py (the reference's own steps on full-frame numpy masks) on a bounded sample of the same
                   workload, one process per host core where the sample has that many independent units
    parity       : the CUDA flow on the CPU sample's inputs == the oracle's kept masks (bit-exact, in order) and measurement rows
"""
import os
import time

import numpy as np

UM_PIX = 0.5
CAP = 100          # Detectron2 TEST.DETECTIONS_PER_IMAGE (not overridden by the reference, src/data/models.py:134-144)


def _params():
    from deepemia_b200 import batched
    # class 0 = large particles, class 1 = small (polyhipes_tommy: pores / throats); manual confidence mode
    return [batched.ClassParams(0, False, 0.5, 0.7), batched.ClassParams(1, True, 0.3, 0.7)]


def _oracle_params():
    return [dict(target_class=0, confidence_threshold=0.5, iou_threshold=0.7), dict(target_class=1, confidence_threshold=0.3, iou_threshold=0.7)]


SMALL = {1}


class TileWorkload:
    """tile_based_inference_pipeline + per-image tail of run_inference on one micrograph (configs 2, 3a, 3b)."""

    def __init__(self, name, h, w, n_particles, tile_size, overlap, upscale, seed, sample):
        self.name, self.h, self.w, self.n_particles = name, h, w, n_particles
        self.tile_size, self.overlap, self.upscale, self.seed = tile_size, overlap, upscale, seed
        self.sample = sample          # dict(h, w, n_particles): the bounded CPU sample (same geometry and density)

    def describe(self):
        return (f"{self.name}: {self.h}x{self.w} micrograph, {self.n_particles} particles, tile {self.tile_size}px overlap "
                f"{self.overlap:.3f} upscale {self.upscale} (full-image pass + tiles, cap 100 heads per predictor call), 2 classes, "
                f"polyhipes rules")

    def generate(self, variant, sample=False):
        from deepemia_b200 import synthetic as syn
        from deepemia_b200 import batched
        g = self.sample if sample else dict(h=self.h, w=self.w, n_particles=self.n_particles)
        d = syn.micrograph_heads(self.seed + variant, g["h"], g["w"], g["n_particles"], self.tile_size, self.overlap, self.upscale)
        # static layout: every predictor call padded to Detectron2's DETECTIONS_PER_IMAGE = 100 (padding = degenerate boxes)
        d["compact"] = dict(full=d["full"], tiles=d["tiles"])
        d["n_real"] = int(d["full"][4][-1] + d["tiles"][4][-1])
        d["full"] = batched.pad_units(*d["full"], CAP)[:5]
        d["tiles"] = batched.pad_units(*d["tiles"], CAP)[:5]
        return d

    @staticmethod
    def heads(data):
        return data["n_real"]

    def to_device(self, data, dev, pinned=False, prob_dtype=np.float32):
        import torch
        from deepemia_b200 import batched
        up = int(self.tile_size * self.upscale)

        def hb(t, H, W):
            conv = lambda a, dt: (torch.as_tensor(np.ascontiguousarray(a.astype(dt))).pin_memory() if pinned else
                                  torch.as_tensor(np.ascontiguousarray(a.astype(dt)), device=dev))
            return batched.HeadBatch(conv(t[0], prob_dtype), conv(t[1], np.float32), conv(t[2], np.float32), conv(t[3], np.int32),
                                     np.asarray(t[4], np.int64), H, W)
        return dict(full=hb(data["full"], *data["hw"]), tiles=hb(data["tiles"], up, up), xy=data["tile_xy"], hw=data["hw"])

    @staticmethod
    def upload(hin, dev):
        from deepemia_b200 import batched
        mv = lambda hb: batched.HeadBatch(hb.probs.to(dev, non_blocking=True), hb.boxes.to(dev, non_blocking=True),
                                          hb.scores.to(dev, non_blocking=True), hb.classes.to(dev, non_blocking=True), hb.unit_off, hb.H, hb.W)
        return dict(full=mv(hin["full"]), tiles=mv(hin["tiles"]), xy=hin["xy"], hw=hin["hw"])

    @staticmethod
    def h2d_bytes(hin):
        return int(sum(t.numel() * t.element_size() for hb in (hin["full"], hin["tiles"]) for t in (hb.probs, hb.boxes, hb.scores, hb.classes)))

    def run(self, din, arena):
        from deepemia_b200 import batched, synthetic as syn
        return batched.tile_pipeline(din["full"], din["tiles"], din["xy"], din["hw"], self.tile_size, self.overlap, _params(), arena,
                                     rules=syn.POLYHIPES_RULES, um_pix=UM_PIX)

    def oracle(self, data):
        """(seconds, heads, masks, classes, rows) of the CPU oracle on `data` (single process: the flow of ONE image is sequential)."""
        from deepemia_b200 import synthetic as syn
        from oracle import flows
        up = int(self.tile_size * self.upscale)
        t0 = time.perf_counter()
        f = data["compact"]["full"]
        full = flows.heads_to_instances(f[0], f[1], f[2], f[3], *data["hw"])
        t = data["compact"]["tiles"]
        tinst = []
        for u in range(len(t[4]) - 1):
            a, b = int(t[4][u]), int(t[4][u + 1])
            tinst.append(flows.heads_to_instances(t[0][a:b], t[1][a:b], t[2][a:b], t[3][a:b], up, up))
        m, s, c = flows.infer_image(full, tinst, [tuple(xy) for xy in data["tile_xy"]], data["hw"], _oracle_params(), SMALL, self.tile_size,
                                    self.overlap, rules=syn.POLYHIPES_RULES)
        rows = flows.measure_rows(m, c, data["hw"], UM_PIX)
        return time.perf_counter() - t0, self.heads(data), m, c, rows


class EnsembleWorkload:
    """config 4: R50 + R101 ensemble merged with multi-scale inference on a batch of images."""
    name = "config4"

    def __init__(self, n_images=64, h=1024, w=1024, n_particles=100, scales=(0.7, 1.0, 1.5), weights=(0.6, 0.4), seed=4000,
                 sample=None):
        self.n_images, self.h, self.w, self.n_particles, self.scales, self.weights, self.seed = n_images, h, w, n_particles, scales, weights, seed
        self.sample = sample or dict(n_images=2, n_particles=24)
        self._compact, self._real = {}, {}

    def describe(self):
        return (f"config4: {self.n_images} images {self.h}x{self.w} x 2 models (weights {self.weights}) x scales {self.scales}, "
                f"{self.n_particles} particles per image (10 % per-model drop-out): universal clean-up -> NEAREST back-projection -> "
                f"score-sorted iou() de-dup 0.4 -> deduplicate_masks_smart -> cross-class 0.7 -> polyhipes rules -> morphometry")

    def generate(self, variant, sample=False):
        from deepemia_b200 import synthetic as syn
        n_img = self.sample["n_images"] if sample else self.n_images
        n_par = self.sample["n_particles"] if sample else self.n_particles
        from deepemia_b200 import batched
        raw = syn.ensemble_multiscale_heads(self.seed + 1000 * variant, n_img, self.h, self.w, n_par, self.scales, len(self.weights))
        d = {s: [batched.pad_units(*t, CAP)[:5] for t in per_model] for s, per_model in raw.items()}
        self._compact[id(d)] = raw
        self._real[id(d)] = int(sum(t[4][-1] for per_model in raw.values() for t in per_model))
        return d

    def heads(self, data):
        return self._real[id(data)]

    def to_device(self, data, dev, pinned=False, prob_dtype=np.float32):
        import torch
        from deepemia_b200 import batched
        conv = lambda a, dt: (torch.as_tensor(np.ascontiguousarray(a.astype(dt))).pin_memory() if pinned else
                              torch.as_tensor(np.ascontiguousarray(a.astype(dt)), device=dev))
        out = {}
        for s, per_model in data.items():
            out[s] = [batched.HeadBatch(conv(t[0], prob_dtype), conv(t[1], np.float32), conv(t[2], np.float32), conv(t[3], np.int32),
                                        np.asarray(t[4], np.int64), int(self.h * s), int(self.w * s)) for t in per_model]
        return out

    @staticmethod
    def upload(hin, dev):
        from deepemia_b200 import batched
        mv = lambda hb: batched.HeadBatch(hb.probs.to(dev, non_blocking=True), hb.boxes.to(dev, non_blocking=True),
                                          hb.scores.to(dev, non_blocking=True), hb.classes.to(dev, non_blocking=True), hb.unit_off, hb.H, hb.W)
        return {s: [mv(hb) for hb in per_model] for s, per_model in hin.items()}

    @staticmethod
    def h2d_bytes(hin):
        return int(sum(t.numel() * t.element_size() for per_model in hin.values() for hb in per_model
                       for t in (hb.probs, hb.boxes, hb.scores, hb.classes)))

    def run(self, din, arena):
        from deepemia_b200 import batched, synthetic as syn
        return batched.ensemble_multiscale(din, list(self.weights), (self.h, self.w), _params(), arena, rules=syn.POLYHIPES_RULES, um_pix=UM_PIX)

    def oracle_image(self, data, b):
        from deepemia_b200 import synthetic as syn
        from oracle import flows
        inst = {}
        for s, per_model in self._compact[id(data)].items():
            inst[s] = []
            for t in per_model:
                a, e = int(t[4][b]), int(t[4][b + 1])
                inst[s].append(flows.heads_to_instances(t[0][a:e], t[1][a:e], t[2][a:e], t[3][a:e], int(self.h * s), int(self.w * s)))
        m, sc, c = flows.ensemble_multiscale(inst, list(self.weights), (self.h, self.w), _oracle_params(), SMALL, rules=syn.POLYHIPES_RULES)
        return m, c, flows.measure_rows(m, c, (self.h, self.w), UM_PIX)

    def oracle(self, data):
        t0 = time.perf_counter()
        B = len(next(iter(self._compact[id(data)].values()))[0][4]) - 1
        res = [self.oracle_image(data, b) for b in range(B)]
        return time.perf_counter() - t0, self.heads(data), res


def _head_batches(din):
    """Every HeadBatch of a workload's device / host input structure."""
    if isinstance(din, dict) and "full" in din:
        return [din["full"], din["tiles"]]
    return [hb for per_model in din.values() for hb in per_model]


def _copy_inputs(dst, src):
    for a, b in zip(_head_batches(dst), _head_batches(src)):
        a.probs.copy_(b.probs, non_blocking=True); a.boxes.copy_(b.boxes, non_blocking=True)
        a.scores.copy_(b.scores, non_blocking=True); a.classes.copy_(b.classes, non_blocking=True)


WORKLOADS = {
    "config2": lambda: TileWorkload("config2", 1024, 1024, 250, 512, 0.1, 2.0, 2000, dict(h=1024, w=1024, n_particles=40)),
    "config3a": lambda: TileWorkload("config3a", 8192, 8192, 20000, 1024, 0.125, 2.0, 3000, dict(h=1100, w=1100, n_particles=22)),
    "config3b": lambda: TileWorkload("config3b", 8192, 8192, 20000, 384, 1.0 / 3.0, 2.0, 3100, dict(h=1024, w=1024, n_particles=80)),
    "config4": lambda: EnsembleWorkload(),
}


def _finish(arena):
    return arena.finish()


def _results_to_host(res, pinned):
    """D2H of what a caller consumes: kept lists + measurement records (capacity-sized) into persistent pinned buffers."""
    out = {}
    for k, v in (("kept_len", res.kept.length), ("kept_idx", res.kept.idx), ("records", res.meas.records), ("rec_inst", res.meas.rec_inst),
                 ("rec_off", res.meas.rec_off), ("scores", res.iset.scores), ("classes", res.iset.classes)):
        buf = pinned.get(k)
        if buf is None or buf.numel() < v.numel() or buf.dtype != v.dtype:
            import torch
            buf = torch.empty(max(int(v.numel() * 1.25), 16), dtype=v.dtype, pin_memory=True)
            pinned[k] = buf
        out[k] = buf[:v.numel()].view(v.shape).copy_(v, non_blocking=True)
    return out


def oracle_leg(name):
    """The CPU oracle on the bounded sample of workload `name` (runs in a worker process of bench.py, one thread)."""
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    import cv2
    import torch
    cv2.setNumThreads(1)
    torch.set_num_threads(1)
    wl = WORKLOADS[name]()
    data = wl.generate(0, sample=True)
    r = wl.oracle(data)
    if isinstance(wl, TileWorkload):
        secs, heads, m, c, rows = r
        per_image = [(np.packbits(np.asarray(m) != 0, axis=None) if len(m) else np.zeros(0, np.uint8), [int(v) for v in c], rows, len(m))]
    else:
        secs, heads, res = r
        per_image = [(np.packbits(np.asarray(m) != 0, axis=None) if len(m) else np.zeros(0, np.uint8), [int(v) for v in c], rows, len(m))
                     for m, c, rows in res]
    return name, secs, heads, per_image


def _parity(wl, dev, leg):
    """The CUDA flow on the CPU sample's inputs against the oracle's result `leg` (from oracle_leg)."""
    import torch
    from deepemia_b200 import engine
    data = wl.generate(0, sample=True)
    din = wl.to_device(data, dev)
    arena = engine.Arena(dev)
    res = None
    for _ in range(8):
        arena.begin()
        res = wl.run(din, arena)
        if arena.finish():
            break
    res.meas.finalize()
    kept = res.kept.to_lists()
    rows = res.meas.rows_to_host()
    cls = res.iset.classes.cpu().numpy()
    ok, detail = True, []
    _, secs, heads, per_image = leg
    hh, ww = (wl.sample["h"], wl.sample["w"]) if isinstance(wl, TileWorkload) else (wl.h, wl.w)
    n_masks = 0
    for b, (bits, c, orows, nm) in enumerate(per_image):
        ids = kept[b]
        n_masks += nm
        if len(ids) != nm:
            ok = False; detail.append(f"image {b}: {len(ids)} kept vs {nm}"); continue
        if ids:
            got = engine.unpack_masks(res.iset, ids).cpu().numpy() != 0
            want = np.unpackbits(bits)[:nm * hh * ww].reshape(nm, hh, ww).astype(bool)
            if not (np.array_equal(got, want) and all(int(cls[ids[i]]) == int(c[i]) for i in range(nm))):
                ok = False; detail.append(f"image {b}: masks differ")
        have = [r[:12] for _, rr in rows[b] for r in rr if r[engine.REC_MEASURED] == 1.0]
        if len(have) != len(orows) or not all(np.allclose(a, np.array(q), rtol=1e-5, atol=0) for a, q in zip(have, orows)):
            ok = False; detail.append(f"image {b}: measurement rows differ")
    return secs, heads, ok, detail, n_masks


def run_workload(name, dev, steps, warmup, variants=2, leg=None):
    """-> dict for bench.py's JSON line."""
    import torch
    from deepemia_b200 import batched, engine
    wl = WORKLOADS[name]()
    datas = [wl.generate(v) for v in range(variants)]
    heads = [wl.heads(d) for d in datas]
    dins = [wl.to_device(d, dev) for d in datas]
    arena = engine.Arena(dev)
    # set-up: eager runs until the arena's capacities cover every variant (these read the totals back), then ONE CUDA graph per
    # resident input set (the padded layouts are identical, the graphs share a memory pool)
    flows = [batched.GraphedFlow((lambda a, d=d: wl.run(d, a)), arena) for d in dins]
    for f in flows:
        f.settle()
    for f in flows:
        f.settle()
    pool = None
    for f in flows:
        f.pool = pool
        f.capture()
        pool = f.graph.pool()
    sticky = torch.zeros(1, dtype=torch.int32, device=dev)
    for i in range(max(warmup, variants)):
        flows[i % variants].replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_heads = 0
    t0 = time.perf_counter()
    e0.record()
    for i in range(steps):
        f = flows[i % variants]
        f.replay()
        sticky.logical_or_(f.abort)          # validity of every step, looked at after the timed region
        n_heads += heads[i % variants]
    e1.record()
    enqueue_ms = (time.perf_counter() - t0) * 1e3
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    valid = not int(sticky.item())
    # ---- end to end from pinned host buffers (probabilities transported as fp16: the reference runs its predictors under AMP
    # autocast by default, src/functions/inference.py:1392-1396; K1 widens them exactly)
    hins = [wl.to_device(d, dev, pinned=True, prob_dtype=np.float16) for d in datas]
    stat = wl.to_device(datas[0], dev, prob_dtype=np.float16)
    ef = batched.GraphedFlow((lambda a: wl.run(stat, a)), arena, pool=pool)
    ef.settle()
    ef.capture()
    pinned = {}

    def e2e_step(i):
        _copy_inputs(stat, hins[i % variants])
        r = ef.replay()
        host = _results_to_host(r, pinned)
        ok = ef.finish()               # the step's one host synchronisation (abort flag + totals, after the result copies)
        return ok, host
    for i in range(max(2, warmup)):
        e2e_step(i)
    torch.cuda.synchronize()
    e0.record()
    ok_all, d2h, n_e2e = True, 0, 0
    for i in range(steps):
        ok, host = e2e_step(i)
        ok_all &= ok
        n_e2e += heads[i % variants]
        d2h = int(sum(v.numel() * v.element_size() for v in host.values()))
    e1.record()
    torch.cuda.synchronize()
    ms_e2e = e0.elapsed_time(e1)
    out = {"workload": wl.describe(), "metric": "instances_per_sec", "unit": "instances/s", "instances_per_step": heads[0],
           "value": n_heads / (ms * 1e-3), "ms_per_step": ms / steps, "host_enqueue_ms_per_step": enqueue_ms / steps,
           "gpu_launches_per_step": flows[0].launches, "cuda_graph": True, "host_syncs_per_step": 0, "input_variants_rotated": variants,
           "all_steps_valid": bool(valid),
           "layout": f"every predictor call padded to {CAP} heads (degenerate boxes, dropped by Boxes.nonempty()): static shapes, one CUDA graph",
           "e2e": {"value": n_e2e / (ms_e2e * 1e-3), "unit": "instances/s", "ms_per_step": ms_e2e / steps,
                   "h2d_bytes_per_step": wl.h2d_bytes(hins[0]), "d2h_bytes_per_step": d2h, "host_syncs_per_step": 1,
                   "probabilities": "fp16 (AMP head output), padded layout", "all_steps_valid": bool(ok_all)},
           "arena_regrow_runs": arena.aborts}
    if leg is not None:
        secs, sheads, ok, detail, n_masks = _parity(wl, dev, leg)
        out["cpu_baseline"] = {"value": sheads / secs, "unit": "instances/s", "cores": 1, "kind": "port",
                               "sample": f"{wl.sample} of the same geometry: {sheads} heads -> {n_masks} final masks in {secs:.1f} s, oracle/flows.py "
                                         "single process (one image's flow is sequential in the reference)"}
        out["parity_vs_gpu_on_sample"] = bool(ok)
        if detail:
            out["parity_detail"] = detail[:4]
    return out


def run_split_workload(name, steps, warmup):
    """One micrograph split over ALL ranks (torchrun; also world size 1): deepemia_b200.distributed.split_micrograph — contiguous
    tile bands, two collectives, identical global stages.  Rank 0 prints the JSON line.  value = heads of the whole micrograph per
    second (max over ranks of the CUDA-event time, barrier on both sides); parity = the split result against the single-GPU batched
    flow on the same micrograph (kept masks bit-exact, in order)."""
    import json
    import torch
    import torch.distributed as dist
    from deepemia_b200 import batched, distributed as D, engine, synthetic as syn
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    wl = WORKLOADS[name]()
    data = wl.generate(0)
    comp = data["compact"]
    T = len(data["tile_xy"])
    t0, t1 = D.band_of_rank(T, rank, world)
    up = int(wl.tile_size * wl.upscale)
    tp = comp["tiles"]
    a, b = int(tp[4][t0]), int(tp[4][t1])
    to = lambda x, dt: torch.as_tensor(np.ascontiguousarray(x.astype(dt)), device=dev)
    tile_hb = batched.HeadBatch(to(tp[0][a:b], np.float32), to(tp[1][a:b], np.float32), to(tp[2][a:b], np.float32), to(tp[3][a:b], np.int32),
                                (np.asarray(tp[4][t0:t1 + 1]) - a).astype(np.int64), up, up) if t1 > t0 else None
    fp = comp["full"]
    full_hb = batched.HeadBatch(to(fp[0], np.float32), to(fp[1], np.float32), to(fp[2], np.float32), to(fp[3], np.int32),
                                np.asarray(fp[4], np.int64), *data["hw"]) if rank == 0 else None
    xy = data["tile_xy"][t0:t1]
    arena = engine.Arena(dev)

    def step():
        return D.split_micrograph(full_hb, tile_hb, xy, data["hw"], wl.tile_size, wl.overlap, _params(), rules=syn.POLYHIPES_RULES,
                                  um_pix=UM_PIX, arena=arena)
    res = None
    for _ in range(max(warmup, 2)):
        res = step()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        res = step()
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # every rank holds the same result: a digest of the final masks, compared across ranks
    kept = res["kept"]
    masks = engine.unpack_masks(res["iset"], kept[:64]).cpu().numpy() if kept else np.zeros((0, 1, 1), np.uint8)
    digest = torch.tensor([len(kept), int(masks.astype(np.int64).sum()), int(np.asarray(kept, np.int64).sum())], dtype=torch.int64, device=dev)
    allg = [torch.empty_like(digest) for _ in range(world)]
    dist.all_gather(allg, digest)
    same_on_all_ranks = all(torch.equal(allg[0], g) for g in allg)
    if rank == 0:
        # single-GPU batched flow on the same micrograph
        din = wl.to_device(data, dev)
        a1 = engine.Arena(dev)
        for _ in range(12):
            a1.begin()
            r1 = wl.run(din, a1)
            if a1.finish():
                break
        k1 = r1.kept.to_lists()[0]
        ok = len(k1) == len(kept)
        for i0 in range(0, min(len(k1), len(kept)), 256):
            m_a = engine.unpack_masks(r1.iset, k1[i0:i0 + 256])
            m_b = engine.unpack_masks(res["iset"], kept[i0:i0 + 256])
            ok = ok and bool(torch.equal(m_a, m_b))
        ms = float(t.item()) / steps
        print(json.dumps({"workload": wl.describe() + f" — split over {world} GPU(s): contiguous tile bands, 2 collectives per step",
                          "metric": "instances_per_sec", "unit": "instances/s", "n_gpus": world, "steps": steps, "warmup": warmup,
                          "value": wl.heads(data) / (ms * 1e-3), "ms_per_step": ms, "instances_per_step": wl.heads(data),
                          "final_masks": len(kept), "collectives_per_step": 2, "host_syncs_per_step": "size read-backs of the exchange (eager flow)",
                          "same_result_on_all_ranks": bool(same_on_all_ranks), "parity_vs_single_gpu_batched_flow": bool(ok)}))
    dist.barrier()
    dist.destroy_process_group()


# ---- row f3: scale-bar line detection for a batch of frames (not a BASELINE config: reported beside them) ---------------------
def _scalebar_cpu(strips, x0):
    """The OpenCV calls of detect_scale_bar (src/utils/scalebar_ocr.py:140, :200, :207-214, :247-249) frame by frame."""
    import cv2
    res = []
    for s in strips:
        gray = cv2.cvtColor(np.ascontiguousarray(s[:, x0:]), cv2.COLOR_BGR2GRAY)
        edges = cv2.Canny(gray, 50, 150, apertureSize=3)
        lines = cv2.HoughLinesP(edges, 1, np.pi / 180, threshold=50, minLineLength=20, maxLineGap=10)
        lines = np.zeros((0, 4), np.int32) if lines is None else lines[:, 0, :]
        sums = []
        for x1, y1, x2, y2 in lines:
            m = np.zeros_like(gray)
            cv2.line(m, (int(x1), int(y1)), (int(x2), int(y2)), 255, 2)
            sums.append((int(gray[m > 0].sum()), int((m > 0).sum())))
        res.append((lines, np.asarray(sums, np.int64).reshape(-1, 2)))
    return res


def run_scalebar_leg(dev, steps, warmup, B=1024, cpu_sample=128):
    import cv2
    import torch
    from deepemia_b200 import engine, synthetic as syn
    rh, W, x0, cap = 92, 1024, 512, 256
    variants = [syn.scalebar_strips(900 + v, B, rh, W, x0) for v in range(2)]
    d_in = [torch.as_tensor(v, device=dev) for v in variants]
    roi = (x0, 0, W, rh)

    def step(t):
        gray, edges = engine.scalebar_edges(t, roi, 50, 150)
        lines, n = engine.hough_lines_p(edges, max_lines=cap)
        return lines, n, engine.line_means(gray, lines, n)
    for i in range(max(warmup, 2)):
        step(d_in[i % 2])
    torch.cuda.synchronize()
    l0 = engine.LAUNCHES["count"]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        res = step(d_in[i % 2])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = (engine.LAUNCHES["count"] - l0) // steps
    # end to end: strips in pinned host memory -> device -> lines, counts and grey sums back on the host
    pin = [torch.as_tensor(v).pin_memory() for v in variants]
    stage = torch.empty_like(d_in[0])
    hres = None

    def e2e(i):
        stage.copy_(pin[i % 2], non_blocking=True)
        lines, n, sums = step(stage)
        out = (lines.cpu(), n.cpu(), sums.cpu())
        return out
    for i in range(2):
        hres = e2e(i)
    torch.cuda.synchronize()
    e0.record()
    for i in range(steps):
        hres = e2e(i)
    e1.record()
    torch.cuda.synchronize()
    ms_e2e = e0.elapsed_time(e1)
    # CPU leg + parity on a sample of the variant the last e2e step used
    v = (steps - 1) % 2
    t0 = time.perf_counter()
    cres = _scalebar_cpu(variants[v][:cpu_sample], x0)
    secs = time.perf_counter() - t0
    lines_h, n_h, sums_h = (a.numpy() for a in hres)
    ok = all(int(n_h[b]) == len(cl) and np.array_equal(lines_h[b, :len(cl)], cl) and np.array_equal(sums_h[b, :len(cl)], cs)
             for b, (cl, cs) in enumerate(cres))
    return {"workload": f"scale-bar line detection (row f3): {B} frames per step, ROI {W - x0}x{rh} px of a 1024x768 frame's info strip: BGR2GRAY -> "
                        "Canny(50,150) -> HoughLinesP(1, pi/180, 50, 20, 10) -> mean grey under every line (thickness-2 mask)",
            "metric": "frames_per_sec", "unit": "frames/s", "value": B * steps / (ms * 1e-3), "ms_per_step": ms / steps,
            "gpu_launches_per_step": int(launches), "mean_lines_per_frame": float(n_h.mean()),
            "e2e": {"value": B * steps / (ms_e2e * 1e-3), "unit": "frames/s", "ms_per_step": ms_e2e / steps,
                    "h2d_bytes_per_step": int(variants[0].nbytes), "d2h_bytes_per_step": int(sum(a.nbytes for a in (lines_h, n_h, sums_h)))},
            "cpu_baseline": {"value": cpu_sample / secs, "unit": "frames/s", "cores": int(cv2.getNumThreads()), "kind": "port",
                             "sample": f"{cpu_sample} of the {B} frames: the reference's own OpenCV calls (cv2 {cv2.__version__}) in one process, "
                                       f"{secs:.2f} s"},
            "parity_vs_gpu_on_sample": bool(ok)}
