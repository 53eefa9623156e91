"""Diagnostic (GPU): why k_contour_measure runs 2x longer per item on a 256-tile shard than on the full sweep.
Counts the work items that take the serial in-thread hull (more than one contour, or more than EMIA_PRESORT_MAX
vertices) and times the morphometry with and without them.   python scripts/measure_tail.py [tiles]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from deepemia_b200 import engine, synthetic as syn  # noqa: E402

tiles = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda", 0)
protos = bench._prototypes()
_, offs, (proto_id, boxes, scores, classes) = bench.shard(0, 1, tiles)
d_probs = torch.as_tensor(protos, device=dev)[torch.as_tensor(proto_id, device=dev)].contiguous()
d_boxes, d_scores, d_classes = (torch.as_tensor(a, device=dev) for a in (boxes, scores, classes))
iset, kept, meas = engine.run_tiles(d_probs, d_boxes, d_scores, d_classes, offs, bench.H, bench.W, um_pix=bench.UM_PIX,
                                    rules=syn.POLYHIPES_RULES, dedup_iou=bench.DEDUP_IOU)
torch.cuda.synchronize()
lists = kept.to_lists()
ids = np.concatenate([np.asarray(l, np.int64) for l in lists])
nc = iset.extra["n_contours"].cpu().numpy()[ids]
cs = iset.cstart.cpu().numpy()[: iset.n * iset.cstart_stride].reshape(iset.n, iset.cstart_stride)
l0 = (cs[ids, 1] - cs[ids, 0])
tot_len = np.array([cs[i, nc_i] - cs[i, 0] for i, nc_i in zip(ids, nc)])
print(f"kept items {len(ids)} of {iset.n}; contours/item: " + ", ".join(f"{k}:{int((nc == k).sum())}" for k in np.unique(nc)))
print("first-contour vertices: mean %.1f  p50 %d  p99 %d  max %d;  > 256: %d;  > 64: %d" % (
    l0.mean(), np.percentile(l0, 50), np.percentile(l0, 99), l0.max(), int((l0 > 256).sum()), int((l0 > 64).sum())))
serial = (nc != 1) | (l0 > 256)
print(f"serial-hull items: {int(serial.sum())}; their total vertices: max {int(tot_len[serial].max()) if serial.any() else 0}")


def timed(groups, reps=5):
    engine.measure_list(iset, groups, um_pix=bench.UM_PIX)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    engine.STAGE_TIMING["enabled"] = True
    engine.stage_summary()
    for _ in range(reps):
        engine.measure_list(iset, groups, um_pix=bench.UM_PIX)
    torch.cuda.synchronize()
    engine.STAGE_TIMING["enabled"] = False
    return engine.stage_summary().get("k5_measure", 0.0)


print("hull + morphometry of the kept lists: %.3f ms" % timed(kept))
# the same lists without the serial-hull items
bad = set(ids[serial].tolist())
lens = kept.length.cpu().numpy().copy()
idx = kept.idx.cpu().numpy().copy()
for g in range(kept.G):
    a = int(kept.cap_off_host[g])
    keep = [i for i in idx[a:a + lens[g]] if int(i) not in bad]
    idx[a:a + len(keep)] = keep
    lens[g] = len(keep)
fast = engine.Groups(cap_off_host=kept.cap_off_host, cap_off=kept.cap_off, length=torch.as_tensor(lens, device=dev),
                     idx=torch.as_tensor(idx, device=dev))
print("  without the serial-hull items     : %.3f ms" % timed(fast))
