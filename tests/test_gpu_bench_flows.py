"""The bench legs of BASELINE configs 2, 3a, 3b and 4 (bench_flows.py) against the CPU oracle of the flows (oracle/flows.py, pinned
to the reference golden by tests/test_oracle_golden.py) on bounded samples of the same geometry — kept masks bit-exact and in
order, classes equal, measurement rows within 1e-5 — and the full-size 8192 x 8192 micrograph of config 3a through the batched
flow against the per-tile mirror (itself pinned by tests/test_gpu_flows.py)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench_flows  # noqa: E402
from deepemia_b200 import batched, engine, synthetic as syn  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["config2", "config3a", "config3b", "config4"])
def test_flow_sample_matches_oracle(cuda_device, name):
    leg = bench_flows.oracle_leg(name)
    wl = bench_flows.WORKLOADS[name]()
    secs, heads, ok, detail, n_masks = bench_flows._parity(wl, cuda_device, leg)
    assert ok, detail
    assert heads > 40 and n_masks > 5


def test_config3a_full_size_batched_equals_per_tile_mirror(cuda_device):
    """The whole 8192 x 8192 micrograph (100 tiles + full-image pass, 8 665 heads, both classes): ONE batched run (sparse K4 on the
    ~4 000-member global lists, CUDA-graph-ready static layout) == the per-tile mirror loop of functions/inference.py."""
    from deepemia_b200.functions import inference as inf
    wl = bench_flows.WORKLOADS["config3a"]()
    data = wl.generate(0)
    din = wl.to_device(data, cuda_device)
    arena = engine.Arena(cuda_device)
    for _ in range(12):
        arena.begin()
        res = wl.run(din, arena)
        if arena.finish():
            break
    kept = res.kept.to_lists()[0]
    assert len(kept) > 2000
    sc = res.iset.scores.cpu().numpy()
    # the mirror: a predictor that replays the synthetic heads of every unit, keyed by the unit's call order
    comp = data["compact"]

    class Replay:
        def __init__(self):
            self.k = 0

        def heads(self, image):
            from deepemia_b200.functions.inference import HeadOutputs
            k = self.k
            self.k += 1
            src, u = (comp["full"], 0) if k % 101 == 0 else (comp["tiles"], k % 101 - 1)
            a, b = int(src[4][u]), int(src[4][u + 1])
            t = lambda x, dt: torch.as_tensor(np.ascontiguousarray(x[a:b].astype(dt)), device=cuda_device)
            return HeadOutputs(t(src[0], np.float32), t(src[1], np.float32), t(src[2], np.float32), t(src[3], np.int64), image.shape[:2])
    image = np.zeros((8192, 8192, 3), np.uint8)
    pred = Replay()
    settings = {"class_0": {"confidence_threshold": 0.5, "iou_threshold": 0.7}, "class_1": {"confidence_threshold": 0.3, "iou_threshold": 0.7}}
    d = inf._infer_image_dev([pred], image, 2, {1}, class_specific_settings=settings, confidence_mode="manual", tile_size=1024,
                             overlap_ratio=0.125, upscale_factor=2.0, spatial_rules=syn.POLYHIPES_RULES, ensemble_enabled=False)
    assert pred.k == 202                                           # (full + 100 tiles) x 2 classes
    assert len(d) == len(kept)
    # 3 000 full-frame masks of 8192 x 8192 would be 190 GB: compare the bit-packed instances through exact fingerprints —
    # bbox, area and the ten raw moments up to order 3 (integer sums of x^p y^q over the mask pixels)
    sel = engine.select(res.iset, kept)
    assert torch.equal(sel.bbox, d.iset.bbox) and torch.equal(sel.area, d.iset.area)
    assert torch.equal(engine.moments(sel)[:, :10], engine.moments(d.iset)[:, :10])
    assert torch.equal(engine.unpack_masks(sel, list(range(8))), engine.unpack_masks(d.iset, list(range(8))))
    assert [float(s) for s in d.scores] == [float(sc[i]) for i in kept]
