"""The per-image visualisation of run_inference (src/functions/inference.py:1080-1145, colours :972-981): every final mask blended
at 50 % in its class colour and outlined, instance number and class name at the centroid.

The blends and outlines of ALL masks run in one kernel on the device (emia_overlay: cv2.addWeighted / cv2.drawContours arithmetic,
in list order, bit-exact); centroids come from the device moments (int(m10 / m00), int(m01 / m00)).  The two anti-aliased
cv2.putText labels per mask are drawn on the host AFTER the kernel — the reference draws them between successive blends, so a
label that a LATER mask overlaps is tinted by that mask in the reference and not here (labels never change a measurement)."""
import cv2
import numpy as np
import torch

from .. import _lib, engine

CLASS_COLORS_BGR = [(0, 255, 0), (255, 0, 0), (0, 0, 255), (255, 255, 0), (255, 0, 255), (0, 255, 255), (128, 0, 128), (255, 165, 0)]


def overlay(image, iset, classes, order=None):
    """uint8 H x W x 3 device tensor: blends + outlines of the instances of `iset` (in `order`, default 0..n-1).  Needs the
    contours of engine.trace() / engine.measure()."""
    lib = _lib.load()
    dev = iset.device
    assert iset.pts is not None, "run engine.trace() / engine.measure() first"
    img = torch.as_tensor(np.ascontiguousarray(image), device=dev).clone() if isinstance(image, np.ndarray) else image.clone()
    assert img.dtype == torch.uint8 and tuple(img.shape) == (iset.H, iset.W, 3)
    cls = torch.as_tensor(np.asarray([int(c) for c in classes], np.int32), device=dev) if not torch.is_tensor(classes) else classes.to(torch.int32)
    colors = torch.as_tensor(np.asarray(CLASS_COLORS_BGR, np.uint8), device=dev)
    order_t = None if order is None else torch.as_tensor(np.asarray(order, np.int32), device=dev)
    n = iset.n if order is None else int(order_t.numel())
    _lib.check(lib.emia_overlay(engine._ptr(img), iset.H, iset.W, engine._ptr(iset.crops), engine._ptr(iset.meta), engine._ptr(iset.crop_off),
                                engine._ptr(order_t), n, engine._ptr(cls), engine._ptr(colors), len(CLASS_COLORS_BGR), engine._ptr(iset.pts),
                                engine._ptr(iset.pt_off), engine._ptr(iset.cstart), iset.cstart_stride, engine._ptr(iset.cont_off),
                                engine._ptr(iset.extra["n_contours"]), engine._stream()), "emia_overlay")
    engine.LAUNCHES["count"] += 1
    return img


def render_predictions(image, iset, classes, thing_classes=None):
    """<name>_predictions.png as a BGR uint8 array."""
    if iset.pts is None:
        engine.trace(iset)
    vis = overlay(image, iset, classes).cpu().numpy()
    mom = engine.moments(iset).cpu().numpy()
    names = list(thing_classes or [])
    for i, cls in enumerate(classes):
        m00, m10, m01 = mom[i, 0], mom[i, 1], mom[i, 2]
        if m00 > 0:
            cX, cY = int(m10 / m00), int(m01 / m00)
            cls = int(cls)
            class_name = names[cls] if cls < len(names) else f"class_{cls}"
            cv2.putText(vis, f"{i + 1}", (cX, cY - 10), cv2.FONT_HERSHEY_SIMPLEX, 0.4, (255, 255, 255), 1, cv2.LINE_AA)
            cv2.putText(vis, class_name, (cX, cY + 15), cv2.FONT_HERSHEY_SIMPLEX, 0.3, (255, 255, 255), 1, cv2.LINE_AA)
    return vis
