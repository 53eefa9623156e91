#!/usr/bin/env python
"""bench.py — throughput of the fused post-head hot path on B200 (BASELINE.json metric: instances measured/sec and
mask MPix/sec; % of HBM roofline).

Workload (BASELINE config 5): 2048 synthetic 1024x1024 tiles x ~Poisson(500) Mask R-CNN head outputs
(28x28 probabilities + box + score + class), sharded tile t -> rank t mod G (strong scaling, no data-path collective).
One step = one pass of the whole path over this rank's shard:
    K1 paste + threshold + bit-pack (full-frame bit masks into a reusable HBM arena + bbox crops + bbox + area)
 -> K5 external contours + morphometry  -> K4 deduplicate_masks_smart(0.7) -> overlap + containment rules.
`value`   : device-resident inputs, CUDA-event timed, max over ranks.
`e2e`     : same path through the public API from pinned HOST buffers, H2D of the step's inputs and D2H of the
            measurement table + kept lists inside the timed region.
`roofline`: dominant kernel (K1) algorithmic bytes / its CUDA-event duration vs MEASURED_PEAKS.json hbm_gbs.
`cpu_baseline`: the CPU oracle (port of the reference's numpy/OpenCV path) on a bounded sample of the same tiles, rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H = W = 1024
N_TILES = 2048
MEAN_INST = 500
UM_PIX = 0.5
DEDUP_IOU = 0.7
N_PROTO = 1024


def _prototypes():
    from deepemia_b200 import synthetic as syn
    rng = np.random.default_rng(12345)
    protos = np.zeros((N_PROTO, 28, 28), np.float32)
    for k in range(N_PROTO):
        poly = syn.star_polygon(rng, 0.0, 0.0, 20.0)
        protos[k], _ = syn.head_from_poly(rng, poly)
    return protos


def tile_heads(t):
    """Head outputs of tile t (seed 5000+t): proto ids, boxes, scores, classes.  ~10 % are jittered re-detections."""
    rng = np.random.default_rng(5000 + t)
    n = max(1, int(rng.poisson(MEAN_INST)))
    proto = rng.integers(0, N_PROTO, n)
    cx = rng.uniform(40, W - 40, n); cy = rng.uniform(40, H - 40, n)
    r = rng.uniform(8, 30, n)
    ax = rng.uniform(0.8, 1.0, n); ay = rng.uniform(0.8, 1.0, n)
    ndup = n // 10
    if ndup:
        src = rng.integers(0, n - ndup, ndup)
        dst = np.arange(n - ndup, n)
        proto[dst] = proto[src]; r[dst] = r[src]; ax[dst] = ax[src]; ay[dst] = ay[src]
        cx[dst] = cx[src] + rng.uniform(-1.5, 1.5, ndup); cy[dst] = cy[src] + rng.uniform(-1.5, 1.5, ndup)
    boxes = np.stack([cx - r * ax, cy - r * ay, cx + r * ax, cy + r * ay], 1).astype(np.float32)
    scores = rng.permutation(n).astype(np.float32)
    scores = (0.05 + 0.95 * (scores + rng.uniform(0.1, 0.9, n).astype(np.float32)) / n).astype(np.float32)
    classes = (rng.random(n) < 0.5).astype(np.int32)
    return proto, boxes, scores, classes


def shard(rank, world, n_tiles):
    tiles = [t for t in range(n_tiles) if t % world == rank]
    parts = [tile_heads(t) for t in tiles]
    offs = np.zeros(len(tiles) + 1, np.int32)
    offs[1:] = np.cumsum([len(p[0]) for p in parts])
    cat = [np.concatenate([p[k] for p in parts]) for k in range(4)]
    return tiles, offs, cat


def _clock_sampler(path):
    q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    try:
        return subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                stdout=open(path, "w"), stderr=subprocess.DEVNULL)
    except Exception:
        return None


def _clock_summary(path, gpu_index):
    sm, mx, reasons = [], [], set()
    try:
        for line in open(path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9 or f[0] != str(gpu_index):
                continue
            sm.append(float(f[1])); mx.append(float(f[2]))
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
    except Exception:
        pass
    if not sm:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
    busy = [s for s in sm if s > 0.5 * max(sm)] or sm
    return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def _oracle_tile(args):
    t, protos = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    import cv2
    import torch
    cv2.setNumThreads(1)
    torch.set_num_threads(1)
    from deepemia_b200 import synthetic as syn
    from oracle import pipeline
    proto, boxes, scores, classes = tile_heads(t)
    final, rows, _ = pipeline.run_tile(protos[proto], boxes, scores, classes, H, W, um_pix=UM_PIX, rules=syn.POLYHIPES_RULES,
                                       dedup_iou=DEDUP_IOU)
    vals = [[float(v) for v in r[3:15]] for r in rows]
    return t, len(proto), final, vals


def cpu_reference_run(tiles, protos, cores):
    """The CPU oracle (port of the reference path) over `tiles`, one process per core.  Returns (seconds, results)."""
    import multiprocessing as mp
    t0 = time.perf_counter()
    if cores > 1:
        with mp.get_context("spawn").Pool(cores) as pool:
            res = pool.map(_oracle_tile, [(t, protos) for t in tiles], chunksize=1)
    else:
        res = [_oracle_tile((t, protos)) for t in tiles]
    return time.perf_counter() - t0, res


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    protos = _prototypes()
    ntl = max(1, min(cores, args.ref_tiles))
    secs, inst = [], 0
    for s in range(args.warmup + args.steps):
        tiles = [(s * ntl + k) % args.tiles for k in range(ntl)]
        dt, res = cpu_reference_run(tiles, protos, min(cores, ntl))
        if s >= args.warmup:
            secs.append(dt); inst += sum(r[1] for r in res)
    total = sum(secs)
    v = inst / total
    line = {"impl": "reference", "metric": "instances_per_sec", "value": v, "unit": "instances/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / max(1, args.steps), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"config5 sample: {ntl} of {args.tiles} tiles (1024x1024, ~500 instances) per step, CPU oracle"},
            "cpu_baseline": {"value": v, "unit": "instances/s", "cores": min(cores, ntl), "kind": "port",
                             "sample": f"{ntl} tiles x {args.steps} steps, one process per core, compute only (no JPEG dump / gc.collect)"},
            "e2e": {"value": v, "unit": "instances/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "mask_mpix_per_sec": v * H * W / 1e6}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--tiles", type=int, default=N_TILES)
    ap.add_argument("--variant", type=int, default=int(os.environ.get("EMIA_PASTE_VARIANT", "2")))
    ap.add_argument("--arena-gb", type=float, default=32.0)
    ap.add_argument("--ref-tiles", type=int, default=8)
    ap.add_argument("--batches", type=int, default=1, help="tile batches of the device-resident pass (1: no overlap; the "
                    "latency-bound kernels lose more under the paste kernel's HBM write stream than the overlap hides)")
    ap.add_argument("--e2e-batches", type=int, default=12, help="tile batches of the three-stream pipeline with host inputs")
    ap.add_argument("--paste-ctas", type=int, default=0, help="resident paste CTAs per SM (0 = kernel default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--breakdown", action="store_true", help="extra untimed pass with per-stage CUDA events (stderr)")
    ap.add_argument("--device-pass-only", action="store_true", help="profiling aid: run only the warm-up + timed device-resident steps")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    from deepemia_b200 import engine, synthetic as syn

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    protos = _prototypes()
    tiles, offs, (proto_id, boxes, scores, classes) = shard(rank, world, args.tiles)
    n_local = int(offs[-1])
    # device-resident inputs
    d_protos = torch.as_tensor(protos, device=dev)
    d_probs = d_protos[torch.as_tensor(proto_id, device=dev)].contiguous()
    d_boxes = torch.as_tensor(boxes, device=dev)
    d_scores = torch.as_tensor(scores, device=dev)
    d_classes = torch.as_tensor(classes, device=dev)
    # pinned host copies for the end-to-end leg
    h_probs = d_probs.cpu().pin_memory(); h_boxes = d_boxes.cpu().pin_memory()
    h_scores = d_scores.cpu().pin_memory(); h_classes = d_classes.cpu().pin_memory()
    pw = engine.pitch_words_for(W)
    frame_bytes = H * pw * 4
    slots = max(1, min(n_local, int(args.arena_gb * 2**30) // frame_bytes))
    arena = torch.empty((slots, H, pw), dtype=torch.int32, device=dev)
    rules = syn.POLYHIPES_RULES

    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    k1_ms = []
    pipe = engine.TilePipeline(H, W, um_pix=UM_PIX, rules=rules, dedup_iou=DEDUP_IOU, frames=arena, variant=args.variant,
                               batches=args.batches, paste_ctas_per_sm=args.paste_ctas, device=dev)
    # host inputs: more batches, so that the H2D copy of batch b+1 hides behind the kernels of batch b
    pipe_e2e = engine.TilePipeline(H, W, um_pix=UM_PIX, rules=rules, dedup_iou=DEDUP_IOU, frames=arena, variant=args.variant,
                                   batches=args.e2e_batches, paste_ctas_per_sm=args.paste_ctas, device=dev)

    def step(probs, bxs, scs, cls, time_k1=False, to_host=False):
        return (pipe_e2e if to_host else pipe).run(probs, bxs, scs, cls, offs, to_host=to_host, time_k1=time_k1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timing ----------------
    res = None
    for _ in range(args.warmup):
        # keep the previous step's results alive while the next one runs, exactly as the timed loop does: the caching
        # allocator then reaches its steady state during the warm-up instead of calling cudaMalloc inside the timed region
        res = step(d_probs, d_boxes, d_scores, d_classes)
    barrier()
    clk_path = os.path.join(ROOT, "gpurun_out", f"clocks_rank{rank}.csv")
    os.makedirs(os.path.dirname(clk_path), exist_ok=True)
    sampler = _clock_sampler(clk_path) if rank == 0 else None
    l0 = engine.LAUNCHES["count"]
    barrier()
    step_ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    k1_events = []
    ev[0].record()
    step_ev[0].record()
    for i in range(args.steps):
        res = step(d_probs, d_boxes, d_scores, d_classes, time_k1=True)
        k1_events.append(list(pipe.k1_events))       # read after the loop: no host synchronisation between steps
        step_ev[i + 1].record()
    ev[1].record()
    barrier()
    k1_ms = [float(sum(a.elapsed_time(b) for a, b in evs)) for evs in k1_events]
    per_step_ms = [step_ev[i].elapsed_time(step_ev[i + 1]) for i in range(args.steps)]
    launches = engine.LAUNCHES["count"] - l0
    ms_total = ev[0].elapsed_time(ev[1])
    assert not pipe.aborted(), "a capacity guard tripped in the timed region: the results of the last step are invalid"
    for r in res:
        r["meas"].finalize()
    if args.breakdown and rank == 0:
        # un-overlapped pass (one batch, one stream) with per-stage CUDA events
        engine.STAGE_TIMING["enabled"] = True
        t0 = time.perf_counter()
        engine.run_tiles(d_probs, d_boxes, d_scores, d_classes, offs, H, W, um_pix=UM_PIX, rules=rules, dedup_iou=DEDUP_IOU,
                         frames=arena, variant=args.variant)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        engine.STAGE_TIMING["enabled"] = False
        summ = engine.stage_summary()
        print(json.dumps({"stage_ms": summ, "sum_ms": sum(summ.values()), "wall_ms": wall}), file=sys.stderr)
    if args.device_pass_only:
        if sampler is not None:
            sampler.terminate()
        if rank == 0:
            print(json.dumps({"device_pass_only": True, "ms_per_step": ms_total / args.steps, "instances": n_local}))
        if world > 1:
            dist.destroy_process_group()
        return
    # K1 alone (nothing else on the GPU), for comparison with its in-pipeline duration
    k1_alone = []
    for _ in range(3):
        ev[2].record()
        iset_alone = engine.paste(d_probs, d_boxes, H, W, frames=arena, variant=args.variant)
        ev[3].record()
        torch.cuda.synchronize()
        k1_alone.append(ev[2].elapsed_time(ev[3]))
    total_crop_words = iset_alone.total_crop_words
    del iset_alone
    # ---------------- end-to-end (host buffers) ----------------
    def e2e_step():
        r = step(h_probs, h_boxes, h_scores, h_classes, to_host=True)
        torch.cuda.current_stream().synchronize()      # the pinned result buffers are complete
        return r
    for _ in range(max(2, args.warmup)):
        res_h = e2e_step()
    barrier()
    e2e_wall = []
    ev[0].record()
    for _ in range(args.steps):
        t0 = time.perf_counter()
        res_h = e2e_step()
        e2e_wall.append(round((time.perf_counter() - t0) * 1e3, 2))
    ev[1].record()
    barrier()
    ms_e2e = ev[0].elapsed_time(ev[1])
    # the same with the head probabilities as fp16 (what the mask head emits under the reference's default AMP autocast;
    # the synthetic probabilities are fp16-representable, K1 widens them exactly): half the H2D bytes
    h_probs16 = h_probs.to(torch.float16).pin_memory()
    assert torch.equal(h_probs16.to(torch.float32), h_probs)
    def e2e16_step():
        r = pipe_e2e.run(h_probs16, h_boxes, h_scores, h_classes, offs, to_host=True)
        torch.cuda.current_stream().synchronize()
        return r
    for _ in range(max(2, args.warmup)):
        res_h16 = e2e16_step()
    barrier()
    ev[0].record()
    for _ in range(args.steps):
        res_h16 = e2e16_step()
    ev[1].record()
    barrier()
    ms_e2e16 = ev[0].elapsed_time(ev[1])
    assert not pipe_e2e.aborted()
    same16 = all(torch.equal(a["host"]["records"], b["host"]["records"]) and torch.equal(a["host"]["kept_idx"], b["host"]["kept_idx"])
                 for a, b in zip(res_h, res_h16))
    if sampler is not None:
        sampler.terminate()
    t = torch.tensor([ms_total, ms_e2e, float(np.mean(k1_ms)), ms_e2e16], dtype=torch.float64, device=dev)
    cnt = torch.tensor([float(n_local)], dtype=torch.float64, device=dev)
    per_rank = [t.clone() for _ in range(world)]
    if world > 1:
        dist.all_gather(per_rank, t)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    per_rank_ms = [round(float(x[0].item()) / args.steps, 3) for x in per_rank]
    ms_total, ms_e2e, k1, ms_e2e16 = t.tolist()
    n_global = cnt.item()
    if rank == 0:
        ms_step = ms_total / args.steps
        value = n_global / (ms_step * 1e-3)
        e2e_val = n_global / (ms_e2e / args.steps * 1e-3)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        traffic = None
        try:     # DRAM bytes per instance of the paste kernel from the committed ncu --set full capture
            tj = json.load(open(os.path.join(ROOT, "profiles", "k1_dram_traffic.json")))
            traffic = float(tj["dram_bytes_per_instance"]) * n_local
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "6650 GB/s (of fallback)"
        crop_bytes = 4.0 * total_crop_words
        k1_bytes = n_local * (28 * 28 * 4 + 16 + 32 + 8 + 16 + 4 + frame_bytes) + crop_bytes   # reads + frame + crop + bbox/area
        k1_gbs = k1_bytes / (k1 * 1e-3) / 1e9
        path_bytes = n_local * (28 * 28 * 4 + 16 + 8 + frame_bytes + 256) + 2 * crop_bytes           # SURVEY §8d B_inst
        h2d = int(h_probs.numel() * 4 + h_boxes.numel() * 4 + h_scores.numel() * 4 + h_classes.numel() * 4)
        d2h = int(sum(v.numel() * v.element_size() for r in res_h for v in r["host"].values()))
        line = {
            "metric": "instances_per_sec", "value": value, "unit": "instances/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"config5: {args.tiles} tiles x ~Poisson(500) instances, 1024x1024, tile t -> rank t mod G",
                       "instances": int(n_global), "paste_variant": args.variant, "frame_arena_gb": round(slots * frame_bytes / 2**30, 1),
                       "tile_batches": args.batches, "e2e_tile_batches": args.e2e_batches,
                       "host_syncs_per_step": 0 if pipe.sync_free else 2, "paste_ctas_per_sm": args.paste_ctas,
                       "per_step_ms": [round(v, 2) for v in per_step_ms], "per_rank_ms_per_step": per_rank_ms,
                       "l2": "per step each rank writes >= 16 GB of frames and re-reads GBs of inputs: far larger than the 126 MB L2",
                       "um_pix": UM_PIX, "dedup_iou": DEDUP_IOU, "rules": "polyhipes_tommy"},
            "mask_mpix_per_sec": value * H * W / 1e6,
            "clocks": _clock_summary(clk_path, local),
            "e2e": {"value": e2e_val, "unit": "instances/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps, "rank0_wall_ms_per_step": e2e_wall},
            "e2e_fp16_heads": {"value": n_global / (ms_e2e16 / args.steps * 1e-3), "unit": "instances/s",
                               "h2d_bytes_per_step": int(h_probs16.numel() * 2 + h_boxes.numel() * 4 + h_scores.numel() * 4 + h_classes.numel() * 4),
                               "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e16 / args.steps, "identical_results_to_f32": bool(same16),
                               "note": "extra: same step with the 28x28 probabilities transported as fp16 (AMP head output), widened exactly in K1"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "k_paste_v2 (K1 paste+threshold+bitpack)", "achieved": k1_gbs, "peak": peak,
                         "unit": "GB/s", "frac": k1_gbs / peak, "traffic": traffic, "algorithmic_bytes": k1_bytes, "peak_source": peak_src,
                         "k1_ms_per_step": k1, "k1_launches_per_step": len(res), "k1_share_of_step": k1 / ms_step,
                         "k1_alone_ms": float(np.median(k1_alone)), "k1_alone_gbs": k1_bytes / (float(np.median(k1_alone)) * 1e-3) / 1e9,
                         "note": "achieved = K1 algorithmic bytes of rank 0 per step / its CUDA-event duration on the stream it is launched on "
                                 "(one launch per tile batch; tile_batches = 1: nothing overlaps it); traffic = dram__bytes_read+write "
                                 "of the same kernel from profiles/k1_dram_traffic.json (ncu --set full), scaled to this step",
                         "path_achieved_gbs": path_bytes / (ms_step * 1e-3) / 1e9,
                         "path_frac": path_bytes / (ms_step * 1e-3) / 1e9 / peak},
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            ntl = max(1, min(len(tiles), args.ref_tiles))          # default 8 tiles (~6 s on 8 cores); more for a wider parity sample
            sample = tiles[:ntl]
            dt, cres = cpu_reference_run(sample, protos, min(cores, ntl))
            inst = sum(r[1] for r in cres)
            # parity of the sampled tiles against the GPU result of the last timed step
            ok = True
            first = res[0]       # the sampled tiles are the first tiles of the shard: all in batch 0
            kl = first["kept"].to_lists(); rows_h = first["meas"].rows_to_host()
            for (tt, n_t, final, vals) in cres:
                g = tiles.index(tt)
                if g >= first["tiles"][1]:
                    continue
                got = [k - int(offs[g]) for k in kl[g]]
                if got != final:
                    ok = False
                    continue
                gv = [r[:12] for _, rr in rows_h[g] for r in rr if r[15] == 1.0]
                if len(gv) != len(vals) or not all(np.allclose(a, np.array(b), rtol=1e-5, atol=0) for a, b in zip(gv, vals)):
                    ok = False
            line["cpu_baseline"] = {"value": inst / dt, "unit": "instances/s", "cores": min(cores, ntl), "kind": "port",
                                    "sample": f"{ntl} of {args.tiles} tiles ({inst} instances), one process per core ({min(cores, ntl)}), {dt:.1f} s, "
                                              "compute only (no per-instance JPEG dump / gc.collect)",
                                    "parity_vs_gpu_on_sample": bool(ok)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
