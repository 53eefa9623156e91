"""Oracle: mask clean-up morphology and RLE (numpy/scipy restatement; test infrastructure only).

Follows the reference:
  * ``rle_encoding``                src/utils/mask_utils.py:17-35   (column-major, 1-indexed start/length pairs)
  * ``postprocess_masks``           src/utils/mask_utils.py:38-84   (Q5 column-sum gate, fill holes, closing with the 3x3 cross,
                                    first-come overlap removal, multi-component masks zeroed but kept — Q6)
  * ``process_masks_parallel``      src/functions/inference.py:170-213 (fill holes, erosion(disk 1), dilation(disk 1))
  * ``postprocess_masks_universal`` src/functions/inference.py:1739-1813
scikit-image 0.19.3 is absent; its ``erosion``/``dilation``/``label`` are thin wrappers over ``scipy.ndimage.grey_erosion`` /
``grey_dilation`` / ``label``, which is what is called here.
"""
import numpy as np
from scipy import ndimage as ndi

CROSS = ndi.generate_binary_structure(2, 1)


def erosion(img):
    return ndi.grey_erosion(img, footprint=CROSS)


def dilation(img):
    return ndi.grey_dilation(img, footprint=CROSS)


def rle_encoding(x):
    dots = np.where(np.asarray(x).T.flatten() == 1)[0]
    runs = []
    prev = -2
    for b in dots:
        if b > prev + 1:
            runs.extend((b + 1, 0))
        runs[-1] += 1
        prev = b
    return runs


def postprocess_masks(ori_mask, ori_score, image_shape, min_crys_size=2):
    height, width = image_shape[:2]
    if len(ori_mask) == 0 or np.asarray(ori_score).all() < 0.5:
        return []
    keep_ind = np.where(np.sum(ori_mask, axis=(0, 1)) > min_crys_size)[0]
    if len(keep_ind) < len(ori_mask):
        if keep_ind.shape[0] != 0:
            ori_mask = ori_mask[: keep_ind.shape[0]]
        else:
            return []
    overlap = np.zeros([height, width])
    out = []
    for i in range(len(ori_mask)):
        mask = ndi.binary_fill_holes(ori_mask[i]).astype(np.uint8)
        mask = erosion(dilation(mask))
        overlap += mask
        mask[overlap > 1] = 0
        lab, _ = ndi.label(mask != 0, structure=ndi.generate_binary_structure(2, 2))
        if lab.max() > 1:
            mask[:] = 0
        out.append(mask)
    return out


def process_masks_parallel(masks):
    return [dilation(erosion(ndi.binary_fill_holes(m).astype(np.uint8))) for m in masks]


def postprocess_masks_universal(ori_mask, image_shape, is_small_class, min_crys_size=None):
    if len(ori_mask) == 0:
        return []
    area = image_shape[0] * image_shape[1]
    if min_crys_size is None:
        min_crys_size = max(3, int(area * 0.000005)) if is_small_class else max(25, int(area * 0.0001))
    out = []
    for m in ori_mask:
        filled = ndi.binary_fill_holes(m).astype(np.uint8)
        final = erosion(filled) if is_small_class else dilation(erosion(filled))
        if np.sum(final) >= min_crys_size:
            out.append(final.astype(bool))
    return out
