"""ORACLE (test infrastructure — never imported by the product path).

numpy restatement of `cv2.connectedComponents[WithStats](img, connectivity=8)` as the reference calls it
(generate_pseudo_label.py:33, engine/runner/loop_UCOD_DPL.py:366,377; OpenCV is an un-vendored dependency,
4.13.0 installed here).  OpenCV's default 8-connectivity labeller (block-based "Spaghetti"/BBDT scan over 2x2
pixel blocks, then label flattening) numbers components in the order in which their first 2x2 block is met in
a raster scan over BLOCKS — i.e. by min over the component's pixels of (y//2)*ceil(W/2) + x//2 — not by first
pixel in pixel-raster order.  Parity pin: tests/test_oracle_golden.py::test_connected_components_match_cv2 compares against cv2 itself on random and
hand-built masks (cv2 is part of the image), and tests/golden/cc_*.npz stores cv2 outputs generated here.
"""
from __future__ import annotations

import numpy as np


def _find(parent, i):
    while parent[i] != i:
        parent[i] = parent[parent[i]]
        i = parent[i]
    return i


def connected_components_8(mask: np.ndarray):
    """mask [H,W] (non-zero = foreground) -> (num_labels, labels int32 [H,W]); label 0 = background,
    labels 1.. in OpenCV order."""
    m = np.asarray(mask) != 0
    H, W = m.shape
    idx = np.arange(H * W, dtype=np.int64).reshape(H, W)
    parent = np.arange(H * W, dtype=np.int64)
    ys, xs = np.nonzero(m)
    # union with the 4 already-visited neighbours (W, NW, N, NE)
    for y, x in zip(ys.tolist(), xs.tolist()):
        a = y * W + x
        for dy, dx in ((0, -1), (-1, -1), (-1, 0), (-1, 1)):
            yy, xx = y + dy, x + dx
            if 0 <= yy < H and 0 <= xx < W and m[yy, xx]:
                ra, rb = _find(parent, a), _find(parent, yy * W + xx)
                if ra != rb:
                    if ra < rb:
                        parent[rb] = ra
                    else:
                        parent[ra] = rb
    roots = np.array([_find(parent, int(i)) for i in idx[m]], dtype=np.int64)
    labels = np.zeros((H, W), dtype=np.int32)
    if roots.size == 0:
        return 1, labels
    bw = (W + 1) // 2
    key = (ys // 2) * bw + (xs // 2)
    uniq, inv = np.unique(roots, return_inverse=True)
    first_key = np.full(uniq.shape, np.iinfo(np.int64).max)
    np.minimum.at(first_key, inv, key)
    order = np.argsort(first_key, kind="stable")
    rank = np.empty_like(order)
    rank[order] = np.arange(1, len(order) + 1)
    labels[ys, xs] = rank[inv]
    return len(uniq) + 1, labels


def stats(labels: np.ndarray, num_labels: int) -> np.ndarray:
    """cv2 CC_STAT_{LEFT,TOP,WIDTH,HEIGHT,AREA} per label (row 0 = background, as OpenCV reports it)."""
    out = np.zeros((num_labels, 5), dtype=np.int32)
    for l in range(num_labels):
        ys, xs = np.nonzero(labels == l)
        if ys.size == 0:
            continue
        out[l] = (xs.min(), ys.min(), xs.max() - xs.min() + 1, ys.max() - ys.min() + 1, ys.size)
    return out


def bounding_rect(binary: np.ndarray):
    """cv2.boundingRect of a binary mask -> (x, y, w, h); (0,0,0,0) if empty."""
    ys, xs = np.nonzero(binary)
    if ys.size == 0:
        return 0, 0, 0, 0
    return int(xs.min()), int(ys.min()), int(xs.max() - xs.min() + 1), int(ys.max() - ys.min() + 1)
