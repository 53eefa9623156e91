"""Contrast d10 / d50 / d90 (src/utils/measurements.py:195-215): the masked 256-bin grey histogram comes from libemia.so
(emia_gray_hist); the percentile read-out over its 256 numbers follows the reference's numpy expressions verbatim."""
import numpy as np

from .. import engine
from . import _bridge

_EDGES = np.linspace(0, 255, 257)


def percentiles_from_counts(counts):
    """counts[256] (np.histogram(pixels, bins=256, range=(0, 255)) puts grey level v into bin v) -> (d10, d50, d90) or Nones."""
    counts = np.asarray(counts, np.float64)
    if counts.sum() <= 0:
        return None, None, None
    hist = counts / np.diff(_EDGES) / counts.sum()          # density=True
    cdf = np.cumsum(hist)
    cdf /= cdf[-1]
    return tuple(np.interp(q, cdf, _EDGES[:-1]) for q in (0.10, 0.50, 0.90))


def contrast_percentiles(original_image, single_im_mask):
    iset = _bridge.upload([np.asarray(single_im_mask) > 0])
    counts = engine.gray_hist(iset, original_image)[0].cpu().numpy()
    return percentiles_from_counts(counts)
