"""CPU container only: time the UNMODIFIED reference functions (imported from /root/reference through tests/golden/refharness.py)
against the oracle port (oracle/) on the SAME tiles of the config-5 sweep, step by step, and record the ratio — bench.py's CPU
baseline runs the port (the reference tree does not exist on the GPU box), this file says how the two compare.

    python scripts/ref_vs_port.py [n_tiles]   ->  profiles/r2_port_vs_reference.json

Steps (after the Detectron2 paste, which both sides take from oracle/d2_paste.py — Detectron2 is absent):
  deduplicate_masks_smart (src/functions/inference.py:2552), filter_by_overlap_rules + filter_by_containment_rules
  (src/utils/spatial_constraints.py:192, :280), findContours + calculate_measurements (inference.py:1164, src/utils/measurements.py:114)."""
import json
import os
import sys
import time

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import bench  # noqa: E402
import refharness  # noqa: E402
from deepemia_b200 import synthetic as syn  # noqa: E402
from oracle import d2_paste, dedup, measure, spatial  # noqa: E402

R = refharness.load_reference()
cv2.setNumThreads(1)
n_tiles = int(sys.argv[1]) if len(sys.argv) > 1 else 8
protos = bench._prototypes()
rules = syn.POLYHIPES_RULES
T = {"port": {"dedup": 0.0, "spatial": 0.0, "measure": 0.0}, "reference": {"dedup": 0.0, "spatial": 0.0, "measure": 0.0}}
inst = same = 0
for t in range(n_tiles):
    proto, boxes, scores, classes = bench.tile_heads(t)
    b, keep = d2_paste.detector_postprocess_boxes(boxes, 1.0, 1.0, bench.H, bench.W)
    masks = [m for m in d2_paste.paste_masks_in_image(protos[proto][keep], b[keep], (bench.H, bench.W))]
    sl = [np.float32(s) for s in scores[keep]]
    cl = [int(c) for c in classes[keep]]
    inst += len(masks)
    out = {}
    for side in ("port", "reference"):
        t0 = time.perf_counter()
        if side == "port":
            m2, s2, c2 = dedup.deduplicate_masks_smart(masks, sl, cl, iou_threshold=0.7)
        else:
            m2, s2, c2 = R.inference.deduplicate_masks_smart(masks, sl, cl, iou_threshold=0.7)
        t1 = time.perf_counter()
        if side == "port":
            m3, s3, c3, _ = spatial.apply_spatial_constraints(m2, s2, c2, rules)
        else:
            m3, s3, c3, _ = R.spatial_constraints.filter_by_overlap_rules(m2, s2, c2, rules["overlap_rules"])
            m3, s3, c3, _ = R.spatial_constraints.filter_by_containment_rules(m3, s3, c3, rules["containment_rules"], rules["containment_threshold"])
        t2 = time.perf_counter()
        rows = []
        if side == "port":
            rows = [[float(v) for v in r[3:15]] for r in measure.measure_masks(m3, c3, (bench.H, bench.W), bench.UM_PIX)]
        else:
            min_area = max(5, bench.H * bench.W * 0.000005 * 0.05)
            for mask in m3:
                binary = (np.asarray(mask) > 0).astype(np.uint8) * 255
                cnts = cv2.findContours(binary, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
                cnts = cnts[0] if len(cnts) == 2 else cnts[1]
                for c in cnts:
                    if cv2.contourArea(c) < min_area:
                        continue
                    m = R.measurements.calculate_measurements(c, binary, um_pix=bench.UM_PIX, pixelsPerMetric=1)
                    rows.append([float(m[k]) for k in measure.MEASUREMENT_KEYS])
        t3 = time.perf_counter()
        T[side]["dedup"] += t1 - t0; T[side]["spatial"] += t2 - t1; T[side]["measure"] += t3 - t2
        out[side] = (len(m3), rows)
    same += int(out["port"][0] == out["reference"][0] and len(out["port"][1]) == len(out["reference"][1]) and
                all(np.allclose(a, q, rtol=0, atol=0) for a, q in zip(out["port"][1], out["reference"][1])))
tot = {k: sum(v.values()) for k, v in T.items()}
res = {"tiles": n_tiles, "instances": inst, "seconds": T, "total_seconds": tot, "port_over_reference_time": tot["port"] / tot["reference"],
       "tiles_with_identical_results": same, "single_process": True,
       "note": "same container, same tiles, single thread; the Detectron2 paste (absent) is taken from oracle/d2_paste.py on both sides and not "
               "timed here; per-instance JPEG dump and gc.collect() of the reference's measurement loop are not run on either side"}
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "profiles", "r2_port_vs_reference.json"), "w"), indent=1)
print(json.dumps(res, indent=1))
