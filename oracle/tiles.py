"""Oracle: tiling helpers (numpy restatement; test infrastructure only).

Follows src/functions/inference.py of the reference: ``generate_tiles_with_overlap`` :2488-2519, ``is_edge_mask`` :2522-2549,
the per-tile back-projection :2399-2420 (cv2.resize INTER_NEAREST, edge filter, placement with clipping)."""
import cv2
import numpy as np


def tile_origins(h, w, tile_size, overlap_ratio):
    stride = int(tile_size * (1 - overlap_ratio))
    return [(x, y) for y in range(0, h, stride) for x in range(0, w, stride)]


def generate_tiles_with_overlap(image, tile_size, overlap_ratio):
    h, w = image.shape[:2]
    out = []
    for x, y in tile_origins(h, w, tile_size, overlap_ratio):
        t = image[y:min(y + tile_size, h), x:min(x + tile_size, w)]
        if t.shape[0] < tile_size or t.shape[1] < tile_size:
            p = np.zeros((tile_size, tile_size, 3), dtype=image.dtype)
            p[:t.shape[0], :t.shape[1]] = t
            t = p
        out.append((t, x, y))
    return out


def is_edge_mask(mask, tile_size, overlap_ratio):
    edge = int(tile_size * overlap_ratio / 2)
    coords = np.argwhere(mask)
    if len(coords) == 0:
        return True
    y_min, x_min = coords.min(axis=0)
    y_max, x_max = coords.max(axis=0)
    return bool(y_min < edge or y_max > tile_size - edge or x_min < edge or x_max > tile_size - edge)


def back_project(mask, tile_w, tile_h, x_offset, y_offset, h, w, tile_size, overlap_ratio, edge_filter_enabled=True):
    """One tile instance -> full-frame bool mask, or None when the edge filter drops it (inference.py:2399-2420)."""
    down = cv2.resize(np.asarray(mask).astype(np.uint8), (tile_w, tile_h), interpolation=cv2.INTER_NEAREST).astype(bool)
    if edge_filter_enabled and is_edge_mask(down, tile_size, overlap_ratio):
        return None
    g = np.zeros((h, w), dtype=bool)
    y_end, x_end = min(y_offset + tile_h, h), min(x_offset + tile_w, w)
    g[y_offset:y_end, x_offset:x_end] = down[:y_end - y_offset, :x_end - x_offset]
    return g
