"""Oracle: Detectron2 0.6 ``detector_postprocess`` + ``paste_masks_in_image`` (CPU path), restated with torch-CPU.

The reference reaches this code through ``predictor(image)`` (src/functions/inference.py:1395,1398,1507,1669,2107; the
predictor is Detectron2's ``DefaultPredictor``, src/data/models.py:107).  Detectron2 (pinned 0.6, requirements.txt:30) is
not vendored under /root/reference and not installed in this image, so its published algorithm is restated here
(SURVEY.md Appendix B.1): ``modeling/postprocessing.py::detector_postprocess`` and ``layers/mask_ops.py::
{paste_masks_in_image,_do_paste_mask}``.  PARITY UNPINNED by reference tests (the reference has none); the arithmetic
itself is ``torch.nn.functional.grid_sample`` of the installed torch (CPU), which is the bit-exactness target.
"""
import numpy as np
import torch
import torch.nn.functional as F


def detector_postprocess_boxes(boxes, scale_x, scale_y, out_h, out_w):
    """Boxes.scale + Boxes.clip + Boxes.nonempty.  Returns (boxes float32 Nx4, keep bool N)."""
    b = torch.as_tensor(np.asarray(boxes, dtype=np.float32)).clone().reshape(-1, 4)
    b[:, 0::2] *= scale_x
    b[:, 1::2] *= scale_y
    b[:, 0].clamp_(min=0, max=out_w)
    b[:, 1].clamp_(min=0, max=out_h)
    b[:, 2].clamp_(min=0, max=out_w)
    b[:, 3].clamp_(min=0, max=out_h)
    keep = ((b[:, 2] - b[:, 0]) > 0) & ((b[:, 3] - b[:, 1]) > 0)
    return b.numpy(), keep.numpy()


def _do_paste_mask(masks, boxes, img_h, img_w):
    # CPU flavour: one instance per call, skip_empty=True
    x0_int, y0_int = torch.clamp(boxes.min(dim=0).values.floor()[:2] - 1, min=0).to(dtype=torch.int32)
    x1_int = torch.clamp(boxes[:, 2].max().ceil() + 1, max=img_w).to(dtype=torch.int32)
    y1_int = torch.clamp(boxes[:, 3].max().ceil() + 1, max=img_h).to(dtype=torch.int32)
    x0, y0, x1, y1 = torch.split(boxes, 1, dim=1)
    N = masks.shape[0]
    img_y = torch.arange(y0_int, y1_int, dtype=torch.float32) + 0.5
    img_x = torch.arange(x0_int, x1_int, dtype=torch.float32) + 0.5
    img_y = (img_y - y0) / (y1 - y0) * 2 - 1
    img_x = (img_x - x0) / (x1 - x0) * 2 - 1
    gx = img_x[:, None, :].expand(N, img_y.size(1), img_x.size(1))
    gy = img_y[:, :, None].expand(N, img_y.size(1), img_x.size(1))
    grid = torch.stack([gx, gy], dim=3)
    img_masks = F.grid_sample(masks, grid.to(masks.dtype), align_corners=False)
    return img_masks[:, 0], (slice(int(y0_int), int(y1_int)), slice(int(x0_int), int(x1_int)))


def paste_masks_in_image(probs, boxes, image_shape, threshold=0.5):
    """probs: N x 28 x 28 float32 probabilities; boxes: N x 4 (already post-processed).  Returns N x H x W bool."""
    probs = torch.as_tensor(np.asarray(probs, dtype=np.float32))
    boxes = torch.as_tensor(np.asarray(boxes, dtype=np.float32)).reshape(-1, 4)
    N = probs.shape[0]
    img_h, img_w = image_shape
    out = torch.zeros(N, img_h, img_w, dtype=torch.bool)
    for i in range(N):
        chunk, sl = _do_paste_mask(probs[i:i + 1, None, :, :], boxes[i:i + 1], img_h, img_w)
        out[(slice(i, i + 1),) + sl] = chunk >= threshold
    return out.numpy()


def predictor_instances(probs, boxes, scores, classes, scale_x, scale_y, out_h, out_w):
    """What ``predictor(image)['instances']`` carries after detector_postprocess: (pred_masks, scores, classes, boxes)."""
    b, keep = detector_postprocess_boxes(boxes, scale_x, scale_y, out_h, out_w)
    probs = np.asarray(probs, dtype=np.float32)[keep]
    b = b[keep]
    masks = paste_masks_in_image(probs, b, (out_h, out_w)) if len(b) else np.zeros((0, out_h, out_w), bool)
    return masks, np.asarray(scores)[keep], np.asarray(classes)[keep], b
