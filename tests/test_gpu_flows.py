"""GPU parity of the reference-interface mirrors (deepemia_b200.functions.inference) against golden vectors produced by the
UNMODIFIED reference functions (tests/golden/make_golden_flows.py: run_class_specific_inference, run_ensemble_inference,
run_iterative_class_inference, process_single_scale, run_adaptive_multiscale_inference, tile_based_inference_pipeline of
src/functions/inference.py, driven by a fake Detectron2 predictor).  Masks, kept sets, order and classes bit-exact; scores equal."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import flow_cases  # noqa: E402

from deepemia_b200 import synthetic as syn  # noqa: E402
from deepemia_b200.functions import inference as inf  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(HERE, "golden", "flows_golden.npz"), allow_pickle=False)


@pytest.mark.parametrize("name", list(flow_cases.CASES))
def test_flow_matches_reference(cuda_device, name):
    case = flow_cases.CASES[name]
    image = flow_cases.make_image(case["image_seed"], *case["shape"])
    preds = [syn.FakeHeadPredictor(**kw) for kw in case["predictors"]]
    predictor = preds if case.get("ensemble") else preds[0]
    old = inf.PARALLEL_MASK_PROCESSING
    inf.PARALLEL_MASK_PROCESSING = case.get("parallel", True)
    try:
        masks, scores, classes = getattr(inf, case["fn"])(predictor, image, *case["args"], **case["kwargs"])
    finally:
        inf.PARALLEL_MASK_PROCESSING = old
    g = {k: GOLD[f"{name}/{k}"] for k in ("bits", "kinds", "scores", "classes", "shape", "empty_arrays")}
    n = len(g["scores"])
    assert len(masks) == n and len(scores) == n and len(classes) == n
    if n == 0:
        assert isinstance(masks, np.ndarray) == bool(g["empty_arrays"])       # ([],[],[]) vs three empty arrays (:1586)
        return
    h, w = (int(v) for v in g["shape"])
    ref = np.unpackbits(g["bits"], axis=1)[:, :h * w].reshape(n, h, w).astype(bool)
    for i in range(n):
        m = np.asarray(masks[i])
        assert m.shape == (h, w) and m.dtype.kind == str(g["kinds"][i]), f"mask {i}: dtype {m.dtype}"
        assert np.array_equal(m != 0, ref[i]), f"mask {i} differs"
        assert float(scores[i]) == float(g["scores"][i]), f"score {i}"
        assert int(classes[i]) == int(g["classes"][i])


@pytest.mark.parametrize("nproc", [1, 2])
def test_split_micrograph(cuda_device, nproc):
    """Config 3 over `nproc` GPUs (2: skipped on a 1-GPU box; the CPU suite covers the two collectives with gloo)."""
    import subprocess
    import torch
    if torch.cuda.device_count() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr", "127.0.0.1",
           "--master-port", str(29611 + nproc), os.path.join(HERE, "dist_split_micrograph.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("OK") == 2 * nproc
