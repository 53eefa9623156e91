"""Mirror of the reference's src/utils/spatial_constraints.py, computing in libemia.so.

Rule sets are plain dicts ({'enabled', 'containment_rules', 'containment_threshold', 'overlap_rules'}), the structure
``load_spatial_constraints`` (:21-67) returns from config/datasets/<name>.yaml; reading that YAML is the reference's config
system (out of scope), so ``apply_spatial_constraints`` accepts the dict directly or a loader callable."""
import numpy as np

from .. import engine
from . import _bridge

_constraint_loader = None


def set_constraint_loader(fn):
    """fn(dataset_name) -> rule dict; the host application plugs its config system in here (load_spatial_constraints)."""
    global _constraint_loader
    _constraint_loader = fn


def get_mask_bbox(mask):
    """(y_min, x_min, y_max, x_max) or None (spatial_constraints.py:70-89)."""
    iset = _bridge.upload([mask])
    b = iset.bbox[0].tolist()
    return None if b[0] < 0 else tuple(b)


def bboxes_overlap(bbox1, bbox2):
    """spatial_constraints.py:92-115 (consistent tuple order, unlike the inference.py variant)."""
    if bbox1 is None or bbox2 is None:
        return False
    y1a, x1a, y1b, x1b = bbox1
    y2a, x2a, y2b, x2b = bbox2
    if x1b < x2a or x2b < x1a:
        return False
    if y1b < y2a or y2b < y1a:
        return False
    return True


def _pair(mask1, mask2):
    iset = _bridge.upload([mask1, mask2])
    inter, a, b = (int(v) for v in engine.pair_counts(iset, [0], [1])[0].tolist())
    return inter, a, b


def calculate_iou(mask1, mask2, bbox1=None, bbox2=None):
    """spatial_constraints.py:118-153."""
    inter, a, b = _pair(mask1, mask2)
    union = a + b - inter
    return inter / union if union > 0 else 0.0


def calculate_containment(child_mask, parent_mask, child_bbox=None, parent_bbox=None):
    """spatial_constraints.py:156-189: fraction of the child's pixels inside the parent."""
    inter, a, _ = _pair(child_mask, parent_mask)
    return inter / a if a > 0 else 0.0


def _removed(before, after):
    return set(before) - set(after)


def filter_by_overlap_rules(masks, scores, classes, overlap_rules):
    """spatial_constraints.py:192-277 -> (masks, scores, classes, removed_indices)."""
    if not masks or not overlap_rules:
        return masks, scores, classes, set()
    iset = _bridge.upload(masks, scores, classes)
    kept = engine.overlap_rules(iset, _bridge.one_group(iset.n, iset.device), overlap_rules).to_lists()[0]
    return [masks[i] for i in kept], [scores[i] for i in kept], [classes[i] for i in kept], _removed(range(len(masks)), kept)


def filter_by_containment_rules(masks, scores, classes, containment_rules, containment_threshold=0.95):
    """spatial_constraints.py:280-398 -> (masks, scores, classes, removed_indices)."""
    if not masks or not containment_rules:
        return masks, scores, classes, set()
    iset = _bridge.upload(masks, scores, classes)
    kept = engine.containment_rules(iset, _bridge.one_group(iset.n, iset.device), containment_rules,
                                    containment_threshold).to_lists()[0]
    return [masks[i] for i in kept], [scores[i] for i in kept], [classes[i] for i in kept], _removed(range(len(masks)), kept)


def apply_spatial_constraints(masks, scores, classes, dataset_name=None, rules=None):
    """spatial_constraints.py:401-460: overlap rules, then containment rules.  `rules` overrides the loader."""
    if rules is None:
        rules = _constraint_loader(dataset_name) if _constraint_loader is not None else None
    if not masks or not rules or not rules.get('enabled', False):
        return masks, scores, classes
    iset = _bridge.upload(masks, scores, classes)
    kept = engine.apply_spatial_constraints(iset, _bridge.one_group(iset.n, iset.device), rules).to_lists()[0]
    return [masks[i] for i in kept], [scores[i] for i in kept], [classes[i] for i in kept]
