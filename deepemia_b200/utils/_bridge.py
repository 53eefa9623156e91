"""List-of-numpy-masks <-> device InstanceSet conversions shared by the reference-interface mirrors.

The reference passes Python lists of H x W numpy arrays (bool or uint8 0/1) between its functions; the
mirrors keep that surface and do all mask arithmetic in libemia.so.  There is no CPU fallback: without CUDA these raise."""
import numpy as np
import torch

from .. import engine, _lib


def device():
    if not torch.cuda.is_available():
        raise _lib.EmiaError("deepemia_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def upload(masks, scores=None, classes=None):
    """list / array of H x W masks (non-zero = set) -> InstanceSet."""
    dev = device()
    arr = np.ascontiguousarray(np.stack([np.asarray(m) != 0 for m in masks]).astype(np.uint8))
    t = torch.as_tensor(arr, device=dev)
    sc = None if scores is None else torch.as_tensor(np.asarray([float(s) for s in scores], np.float32), device=dev)
    cl = None if classes is None else torch.as_tensor(np.asarray([int(c) for c in classes], np.int32), device=dev)
    return engine.from_masks(t, scores=sc, classes=cl)


def download(iset, idx=None, dtype=np.uint8):
    """Full-frame numpy masks (list) of the selected instances."""
    if idx is not None and len(idx) == 0:
        return []
    out = engine.unpack_masks(iset, idx).cpu().numpy()
    if dtype is bool:
        out = out.astype(bool)
    return [m for m in out]


def one_group(n, dev):
    return engine.groups_from_offsets([0, n], dev)


def list_group(idx, dev):
    return engine.groups_from_lists([list(idx)], dev)
