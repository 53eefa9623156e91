"""Mirror of the reference's src/utils/measurements.py::calculate_measurements (:114-233), computing in libemia.so.

Value types follow the reference as it executes under numpy 2.x: the quantities derived from
``order_points`` (float32) are np.float32, the ellipse axes / Circularity / Chords Python floats, the rest np.float64."""
import numpy as np

from .. import engine

_F32 = ("Length", "Width", "Aspect_Ratio", "Roundness", "Feret_diam")
_PYFLOAT = ("major_axis_length", "minor_axis_length", "Circularity", "Chords")
_FIELDS = {"major_axis_length": engine.REC_MAJOR, "minor_axis_length": engine.REC_MINOR, "eccentricity": engine.REC_ECC,
           "Length": engine.REC_LENGTH, "Width": engine.REC_WIDTH, "CircularED": engine.REC_CED, "Aspect_Ratio": engine.REC_ASPECT,
           "Circularity": engine.REC_CIRC, "Chords": engine.REC_CHORDS, "Feret_diam": engine.REC_FERET, "Roundness": engine.REC_ROUND,
           "Sphericity": engine.REC_SPHER}
KEY_ORDER = ["major_axis_length", "minor_axis_length", "eccentricity", "Length", "Width", "CircularED", "Aspect_Ratio", "Circularity",
             "Chords", "Feret_diam", "Roundness", "Sphericity", "contrast_d10", "contrast_d50", "contrast_d90"]


def record_to_dict(rec, n_vertices=None):
    """One 16-double record (engine.REC_* order) -> the reference's measurement dict (15 keys, :217-233)."""
    out = {}
    for k, f in _FIELDS.items():
        v = float(rec[f])
        if k in _F32:
            out[k] = np.float32(v)
        elif k in _PYFLOAT:
            out[k] = v
        else:
            out[k] = np.float64(v)
    nv = int(rec[engine.REC_NVERT]) if n_vertices is None else n_vertices
    if nv < 5:                                    # no ellipse fit below 5 points (:176): the reference stores int 0
        out["major_axis_length"] = out["minor_axis_length"] = out["eccentricity"] = 0
    out["contrast_d10"] = out["contrast_d50"] = out["contrast_d90"] = None
    return {k: out[k] for k in KEY_ORDER}


def calculate_measurements(c, single_im_mask, um_pix=1.0, pixelsPerMetric=1.0, original_image=None,
                           measure_contrast_distribution=False):
    """calculate_measurements (src/utils/measurements.py:114-233) for one OpenCV contour `c` ([K,1,2] int32)."""
    if pixelsPerMetric != 1 and pixelsPerMetric != 1.0:
        raise ValueError("pixelsPerMetric is always 1 in the reference (src/functions/inference.py:1174); other values are unsupported")
    rec = engine.measure_contours([np.asarray(c).reshape(-1, 2)], um_pix=um_pix)[0].cpu().numpy()
    out = record_to_dict(rec, n_vertices=len(c))
    if measure_contrast_distribution and original_image is not None:
        from .contrast import contrast_percentiles
        out["contrast_d10"], out["contrast_d50"], out["contrast_d90"] = contrast_percentiles(original_image, single_im_mask)
    return out
