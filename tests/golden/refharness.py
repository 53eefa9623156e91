"""Import the UNMODIFIED reference (Deam0on/deepEMIA) from /root/reference in this container.

Test infrastructure only: used by ``make_golden.py`` to generate the committed golden vectors.
/root/reference does not exist on the GPU box, so nothing under ``-m gpu``, ``smoke()`` or
``bench.py`` may import this module.

Recipe: HOME redirected so that ``~/deepEMIA/config`` resolves to the reference's
own ``config/`` directory (``src/utils/config.py:82`` hard-wires ``Path.home()``), MagicMock stubs for
the absent heavyweight third-party packages (detectron2, easyocr, shapely), and *functional* shims
for the two small absent packages whose arithmetic is on the path:

* ``skimage.morphology.{disk,erosion,dilation}``, ``skimage.measure.label`` — scikit-image 0.19.3
  implements them as thin wrappers over ``scipy.ndimage.grey_erosion / grey_dilation / label``.
* ``imutils.grab_contours``, ``imutils.perspective.order_points``.
"""
import os
import sys
import tempfile
import types
from unittest.mock import MagicMock

REFERENCE_ROOT = "/root/reference"
_loaded = {}


def _skimage_shim():
    import numpy as np
    from scipy import ndimage as ndi

    sk = types.ModuleType("skimage")
    morph = types.ModuleType("skimage.morphology")
    meas = types.ModuleType("skimage.measure")

    def disk(radius, dtype=np.uint8):
        L = np.arange(-radius, radius + 1)
        X, Y = np.meshgrid(L, L)
        return np.array((X ** 2 + Y ** 2) <= radius ** 2, dtype=dtype)

    def _fp(image, footprint):
        if footprint is None:
            footprint = ndi.generate_binary_structure(image.ndim, 1)
        return np.asarray(footprint, dtype=bool)

    def erosion(image, footprint=None, out=None, shift_x=False, shift_y=False):
        # skimage 0.19.3 morphology/grey.py: ndi.grey_erosion(image, footprint=footprint, output=out)
        fp = _fp(image, footprint)
        if out is None:
            out = np.empty_like(image)
        ndi.grey_erosion(image, footprint=fp, output=out)
        return out

    def dilation(image, footprint=None, out=None, shift_x=False, shift_y=False):
        # skimage 0.19.3 inverts the footprint before grey_dilation; symmetric footprints are unaffected
        fp = _fp(image, footprint)
        fp = fp[::-1, ::-1]
        if out is None:
            out = np.empty_like(image)
        ndi.grey_dilation(image, footprint=fp, output=out)
        return out

    def label(label_image, background=None, return_num=False, connectivity=None):
        # default connectivity = ndim (8-connected in 2-D)
        if connectivity is None:
            connectivity = label_image.ndim
        st = ndi.generate_binary_structure(label_image.ndim, connectivity)
        lab, num = ndi.label(label_image != 0, structure=st)
        return (lab, num) if return_num else lab

    morph.disk, morph.erosion, morph.dilation = disk, erosion, dilation
    meas.label = label
    sk.morphology, sk.measure = morph, meas
    return {"skimage": sk, "skimage.morphology": morph, "skimage.measure": meas}


def _imutils_shim():
    import numpy as np

    im = types.ModuleType("imutils")
    persp = types.ModuleType("imutils.perspective")

    def grab_contours(cnts):
        if len(cnts) == 2:
            return cnts[0]
        if len(cnts) == 3:
            return cnts[1]
        raise Exception("Contours tuple must have length 2 or 3")

    def order_points(pts):
        from scipy.spatial import distance as dist
        xSorted = pts[np.argsort(pts[:, 0]), :]
        leftMost = xSorted[:2, :]
        rightMost = xSorted[2:, :]
        leftMost = leftMost[np.argsort(leftMost[:, 1]), :]
        (tl, bl) = leftMost
        D = dist.cdist(tl[np.newaxis], rightMost, "euclidean")[0]
        (br, tr) = rightMost[np.argsort(D)[::-1], :]
        return np.array([tl, tr, br, bl], dtype="float32")

    im.grab_contours = grab_contours
    persp.order_points = order_points
    im.perspective = persp
    return {"imutils": im, "imutils.perspective": persp}


def load_reference():
    """Returns a namespace with the reference modules (inference, mask_utils, spatial_constraints, measurements)."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError("reference tree not present (only available in the build container)")
    home = tempfile.mkdtemp(prefix="emia_refhome_")
    os.makedirs(os.path.join(home, "deepEMIA"))
    os.symlink(os.path.join(REFERENCE_ROOT, "config"), os.path.join(home, "deepEMIA", "config"))
    os.environ["HOME"] = home
    for name in [
        "detectron2", "detectron2.data", "detectron2.data.transforms", "detectron2.utils",
        "detectron2.utils.visualizer", "detectron2.config", "detectron2.engine", "detectron2.data.datasets",
        "detectron2.structures", "detectron2.model_zoo", "detectron2.evaluation", "easyocr",
        "shapely", "shapely.affinity", "shapely.geometry",
    ]:
        sys.modules.setdefault(name, MagicMock())
    for k, v in {**_skimage_shim(), **_imutils_shim()}.items():
        sys.modules.setdefault(k, v)
    sys.path.insert(0, REFERENCE_ROOT)
    import logging
    from src.utils import mask_utils, measurements, spatial_constraints
    from src.functions import inference
    logging.getLogger("system").setLevel(logging.ERROR)
    for h in logging.getLogger("system").handlers:
        h.setLevel(logging.ERROR)
    _loaded.update(inference=inference, mask_utils=mask_utils,
                   spatial_constraints=spatial_constraints, measurements=measurements)
    return types.SimpleNamespace(**_loaded)
