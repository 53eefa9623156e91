"""BASELINE config 2: a random-init torchvision Mask R-CNN R50-FPN on one 1024 x 1024 image — Detectron2 is absent,
torchvision's R50-FPN has the same 28 x 28 mask head.  The pre-paste probabilities are captured by wrapping
torchvision.models.detection.roi_heads.maskrcnn_inference (torchvision's own paste is NOT the oracle); probs + boxes go through
K1 -> contours -> de-dup -> spatial constraints -> morphometry and must match the oracle (Detectron2-paste restatement + the
reference's own steps) bit-exactly / within 1e-5."""
import numpy as np
import pytest
import torch

from deepemia_b200 import engine, synthetic as syn
from oracle import d2_paste, pipeline

pytestmark = pytest.mark.gpu


def _torchvision_heads(device):
    tv = pytest.importorskip("torchvision")
    import torchvision.models.detection.roi_heads as rh
    from torchvision.models.detection import maskrcnn_resnet50_fpn
    torch.manual_seed(0)
    model = maskrcnn_resnet50_fpn(weights=None, weights_backbone=None, box_score_thresh=0.0, num_classes=3).eval().to(device)
    captured = {}
    orig = rh.maskrcnn_inference

    def spy(x, labels):
        out = orig(x, labels)
        captured["probs"] = out[0][:, 0].detach().float().contiguous()
        return out

    rh.maskrcnn_inference = spy
    try:
        g = torch.Generator().manual_seed(0)
        image = torch.rand(3, 1024, 1024, generator=g).to(device)
        with torch.no_grad():
            det = model([image])[0]
    finally:
        rh.maskrcnn_inference = orig
    n = det["boxes"].shape[0]
    assert n > 0 and captured["probs"].shape == (n, 28, 28)
    return captured["probs"], det["boxes"].float().contiguous(), det["scores"].float().contiguous()


def test_config2_single_image(cuda_device):
    H = W = 1024
    probs, boxes, scores = _torchvision_heads(cuda_device)
    n = probs.shape[0]
    # random init gives one dominant label: classes = index mod 2 so that the spatial rules have work;
    # detection scores of a random-init head collide, the path is specified for distinct scores
    classes = torch.arange(n, device=cuda_device, dtype=torch.int32) % 2
    scores = (scores + torch.linspace(0, 1e-3, n, device=cuda_device)).float().contiguous()
    assert len(torch.unique(scores)) == n
    # paste parity on real head outputs (box sizes from a few to hundreds of pixels)
    ref_masks, _, _, _ = d2_paste.predictor_instances(probs.cpu().numpy(), boxes.cpu().numpy(), scores.cpu().numpy(), classes.cpu().numpy(),
                                                      1.0, 1.0, H, W)
    iset = engine.paste(probs, boxes, H, W, frames=True)
    keep = iset.valid.cpu().numpy()
    got = engine.unpack_masks(iset).cpu().numpy().astype(bool)[keep]
    assert got.shape == ref_masks.shape and np.array_equal(got, ref_masks)
    # the fused path
    iset, kept, meas = engine.run_tiles(probs, boxes, scores, classes, [0, n], H, W, um_pix=0.5, rules=syn.POLYHIPES_RULES, dedup_iou=0.7)
    final, rows, _ = pipeline.run_tile(probs.cpu().numpy(), boxes.cpu().numpy(), scores.cpu().numpy(), classes.cpu().numpy(), H, W,
                                       um_pix=0.5, rules=syn.POLYHIPES_RULES, dedup_iou=0.7)
    assert kept.to_lists()[0] == final
    vals = [r[:12] for _, rr in meas.rows_to_host()[0] for r in rr if r[engine.REC_MEASURED] == 1.0]
    assert len(vals) == len(rows)
    undefined = 0
    degenerate = sum(1 for r in rows if float(r[3 + 3]) <= 0.5 + 1e-9)          # C. Length = short side of minAreaRect x um_pix: <= 1 px
    for g, r in zip(vals, rows):
        ref = np.array([float(v) for v in r[3:15]])
        if np.allclose(g, ref, rtol=1e-5, atol=1e-12):
            continue
        # The only tolerated disagreement: the three ellipse columns of a strip two pixels wide (minAreaRect short side 1 px).
        # Its vertices lie on two parallel lines, a degenerate conic on which cv2.fitEllipse returns NaN, a rounding-noise axis
        # or — re-fitting RANDOMLY perturbed points — a different answer on every call (scripts/debug_cfg2.py).
        assert ref[3] <= 0.5 + 1e-9, "ellipse columns differ on a non-degenerate contour"
        np.testing.assert_allclose(g[3:], ref[3:], rtol=1e-5, atol=1e-12)
        undefined += 1
    print(f"PARITY-COUNT config2: rows {len(rows)} degenerate-strips {degenerate} ellipse-undefined {undefined}")
    assert undefined <= degenerate, (undefined, degenerate, len(rows))         # only degenerate strips may differ, nothing else
