"""Golden vectors for the inference FLOWS: the UNMODIFIED reference functions (imported from /root/reference through
refharness.py) driven by a fake Detectron2 predictor whose `instances` are the Detectron2-paste oracle (oracle/d2_paste.py)
of deterministic synthetic head outputs (deepemia_b200.synthetic.FakeHeadPredictor).  Run once in the build container:

    python tests/golden/make_golden_flows.py        ->  tests/golden/flows_golden.npz

Recorded per case: the returned masks (np.packbits), their dtype kinds, scores (float64) and classes.
The same cases are replayed on the GPU through deepemia_b200.functions.inference in tests/test_gpu_flows.py."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import refharness  # noqa: E402
from deepemia_b200 import synthetic as syn  # noqa: E402
from oracle import d2_paste  # noqa: E402
import flow_cases  # noqa: E402

R = refharness.load_reference()
inf = R.inference


class _Field:
    def __init__(self, a):
        self.a = a

    def cpu(self):
        return self

    def numpy(self):
        return self.a


class FakeInstances:
    """What the reference touches of detectron2.structures.Instances (inference.py:1401-1403, :1509-1516)."""

    def __init__(self, masks, scores, classes):
        self._fields = {"pred_masks": _Field(masks), "scores": _Field(scores), "pred_classes": _Field(classes)}
        self.pred_masks, self.scores, self.pred_classes = (self._fields[k] for k in ("pred_masks", "scores", "pred_classes"))

    def to(self, _):
        return self

    def __len__(self):
        return len(self._fields["scores"].a)


class RefPredictor:
    """DefaultPredictor stand-in: heads -> detector_postprocess + paste_masks_in_image (oracle) -> Instances."""

    def __init__(self, fake):
        self.fake = fake

    def __call__(self, image):
        probs, boxes, scores, classes, (in_h, in_w) = self.fake.raw_heads(image)
        H, W = image.shape[:2]
        masks, s, c, _ = d2_paste.predictor_instances(probs, boxes, scores, classes, W / in_w, H / in_h, H, W)
        return {"instances": FakeInstances(masks, s.astype(np.float32), c.astype(np.int64))}


def pack(masks, scores, classes):
    masks = list(masks)
    if len(masks) == 0:
        return dict(bits=np.zeros((0, 0), np.uint8), kinds=np.zeros(0, "U1"), scores=np.zeros(0), classes=np.zeros(0, np.int64),
                    shape=np.zeros(2, np.int64), empty_type=np.array(type(masks).__name__ if not isinstance(masks, np.ndarray) else "ndarray"))
    arr = np.stack([np.asarray(m) != 0 for m in masks])
    return dict(bits=np.packbits(arr.reshape(len(masks), -1), axis=1), kinds=np.array([np.asarray(m).dtype.kind for m in masks]),
                scores=np.array([float(s) for s in scores], np.float64), classes=np.array([int(c) for c in classes], np.int64),
                shape=np.array(arr.shape[1:], np.int64), empty_type=np.array("list"))


def main():
    out = {}
    for name, case in flow_cases.CASES.items():
        image = flow_cases.make_image(case["image_seed"], *case["shape"])
        fakes = [syn.FakeHeadPredictor(**kw) for kw in case["predictors"]]
        preds = [RefPredictor(f) for f in fakes]
        predictor = preds if case.get("ensemble") else preds[0]
        inf.PARALLEL_MASK_PROCESSING = case.get("parallel", True)
        fn = getattr(inf, case["fn"])
        res = fn(predictor, image, *case["args"], **case["kwargs"])
        for k, v in pack(*res).items():
            out[f"{name}/{k}"] = v
        ret_empty_arrays = isinstance(res[0], np.ndarray) and len(res[0]) == 0
        out[f"{name}/empty_arrays"] = np.array(ret_empty_arrays)
        print(name, "->", len(res[0]), "masks,", sum(f.calls for f in fakes), "predictor calls")
    np.savez_compressed(os.path.join(HERE, "flows_golden.npz"), **out)


if __name__ == "__main__":
    main()
