import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, cv2, torch
from scipy import ndimage as ndi
from deepemia_b200 import engine, synthetic as syn
from oracle import morphology
H, W = 160, 200
rng = np.random.default_rng(2)
polys = syn.particle_field(rng, 14, H, W, rmin=6, rmax=22, margin=10)
polys += [p + rng.uniform(-2.5, 2.5, 2) for p in polys[:8]]
ms = [m.astype(bool) for m in syn.masks_from_polys(polys, H, W)]
ring = np.zeros((H, W), np.uint8); cv2.circle(ring, (90, 80), 25, 1, 3); ms.insert(2, ring.astype(bool))
dev = torch.device("cuda:0")
iset = engine.from_masks(torch.as_tensor(np.stack(ms).astype(np.uint8), device=dev))
closed = engine.morph(iset, [engine.MORPH_FILL, engine.MORPH_DILATE, engine.MORPH_ERODE])
gc = engine.unpack_masks(closed).cpu().numpy()
for i, m in enumerate(ms):
    r = morphology.erosion(morphology.dilation(ndi.binary_fill_holes(m).astype(np.uint8)))
    if not np.array_equal(gc[i], r):
        print("closing differs", i, int((gc[i] != r).sum()), np.argwhere(gc[i] != r)[:5], "bbox", iset.bbox[i].tolist())
ref = morphology.postprocess_masks(np.stack(ms), np.ones(len(ms), np.float32), (H, W), min_crys_size=2)
out, gated = engine.postprocess_masks(iset, engine.groups_from_offsets([0, len(ms)], dev), 2)
go = engine.unpack_masks(out).cpu().numpy()
for i, r in enumerate(ref):
    if not np.array_equal(go[i], r):
        print("final differs", i, "gpu sum", int(go[i].sum()), "ref sum", int(r.sum()), np.argwhere(go[i] != r)[:5])
