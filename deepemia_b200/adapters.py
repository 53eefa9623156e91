"""Pre-paste head adapters: objects with

    heads(image) -> HeadOutputs(probs [N,28,28], boxes [N,4] xyxy in MODEL-INPUT coordinates, scores [N], classes [N], (in_h, in_w))

i.e. the detector run WITHOUT its own paste, so that detector_postprocess + paste_masks_in_image (src/functions/inference.py:
1395, 1398, 1507, 1669, 2107 via predictor(image); predictor built at src/data/models.py:107) happen in K1 and no full-frame
mask ever crosses PCIe.  A list of adapters is an ensemble, exactly like a list of predictors in the reference."""
import numpy as np
import torch

from .functions.inference import HeadOutputs


class Detectron2HeadAdapter:
    """Wraps a detectron2.engine.DefaultPredictor (the object src/data/models.py:107 returns): replicates its preprocessing
    (BGR/RGB flip, ResizeShortestEdge(800, 1333)) and calls GeneralizedRCNN.inference(do_postprocess=False).
    use_amp mirrors the reference's `torch.cuda.amp.autocast()` around predictor(image) (inference.py:1392-1396): the mask head
    then emits fp16 probabilities, which K1 reads directly."""

    def __init__(self, predictor, use_amp=True):
        self.p = predictor
        self.use_amp = use_amp

    def heads(self, image):
        p = self.p
        with torch.no_grad():
            im = image[:, :, ::-1] if p.input_format == "RGB" else image
            t = p.aug.get_transform(im).apply_image(im)
            x = {"image": torch.as_tensor(np.ascontiguousarray(t.astype("float32").transpose(2, 0, 1))), "height": image.shape[0],
                 "width": image.shape[1]}
            with torch.autocast("cuda", enabled=self.use_amp and torch.cuda.is_available()):
                inst = p.model.inference([x], do_postprocess=False)[0]
        return HeadOutputs(inst.pred_masks[:, 0], inst.pred_boxes.tensor.float(), inst.scores.float(), inst.pred_classes,
                           tuple(int(v) for v in t.shape[:2]))


class TorchvisionHeadAdapter:
    """torchvision.models.detection.MaskRCNN (same 28 x 28 mask head as Detectron2's R50/R101-FPN): the model's forward without
    transform.postprocess — roi_heads already returns the per-detection sigmoid probabilities [N, 1, 28, 28] and the boxes in
    the RESIZED input's coordinates; torchvision's own paste (expand_boxes + interpolate) is skipped.  Labels are 1-based in
    torchvision (0 = background): class = label - 1, as a Detectron2 model trained on the same categories would number them."""

    def __init__(self, model, device=None, use_amp=False):
        self.model = model.eval()
        self.device = torch.device(device) if device is not None else next(model.parameters()).device
        self.use_amp = use_amp

    def heads(self, image):
        """image: H x W x 3 BGR uint8 (what the reference hands to predictor(image)) or a float tensor [3, H, W] in 0..1."""
        m = self.model
        if isinstance(image, np.ndarray):
            x = torch.as_tensor(np.ascontiguousarray(image[:, :, ::-1].transpose(2, 0, 1))).to(self.device).float() / 255.0
        else:
            x = image.to(self.device)
        with torch.no_grad(), torch.autocast("cuda", enabled=self.use_amp and self.device.type == "cuda"):
            images, _ = m.transform([x])
            features = m.backbone(images.tensors)
            if isinstance(features, torch.Tensor):
                features = {"0": features}
            proposals, _ = m.rpn(images, features)
            det, _ = m.roi_heads(features, proposals, images.image_sizes)
        d = det[0]
        n = d["boxes"].shape[0]
        probs = d["masks"][:, 0] if "masks" in d and n else torch.zeros((0, 28, 28), device=self.device)
        return HeadOutputs(probs.contiguous(), d["boxes"].float().contiguous(), d["scores"].float().contiguous(),
                           (d["labels"] - 1).to(torch.int32).contiguous(), tuple(int(v) for v in images.image_sizes[0]))


def calculate_average_mask_sizes(predictors, images_sample, metadata=None, read_image=None):
    """calculate_average_mask_sizes (src/functions/inference.py:1626-1705): {class_id: mean mask area} over the detections with
    score >= 0.7 of (at most) the first 5 sample images, first predictor only.  The areas are K1's popcounts (no mask leaves the
    device).  images_sample: paths (read with cv2.imread, as the reference) or arrays."""
    import cv2
    from . import engine
    class_sizes = {}
    for item in list(images_sample)[:min(5, len(images_sample))]:
        image = (read_image or cv2.imread)(item) if isinstance(item, str) else item
        if image is None:
            continue
        ho = predictors[0].heads(image)
        H, W = image.shape[:2]
        in_h, in_w = ho.input_size
        probs = ho.probs.reshape(-1, engine.MASK_SIDE, engine.MASK_SIDE)
        if probs.dtype not in (torch.float16, torch.float32):
            probs = probs.float()
        iset = engine.paste(probs.contiguous(), ho.boxes.float().contiguous(), H, W, scale_x=float(W) / in_w, scale_y=float(H) / in_h)
        keep = (iset.valid & (ho.scores.float() >= 0.7)).cpu().numpy()
        area = iset.area.cpu().numpy()
        cls = ho.classes.cpu().numpy()
        for a, c, k in zip(area, cls, keep):
            if k:
                class_sizes.setdefault(int(c), []).append(int(a))
    return {c: np.mean(v) for c, v in class_sizes.items() if v}


def determine_small_classes(class_avg_sizes, threshold_percentile=50):
    """determine_small_classes (src/functions/inference.py:1708-1736): classes whose mean mask size is <= the percentile."""
    if not class_avg_sizes:
        return set()
    threshold_size = np.percentile(list(class_avg_sizes.values()), threshold_percentile)
    return {cls for cls, size in class_avg_sizes.items() if size <= threshold_size}
