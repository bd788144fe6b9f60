"""ORACLE (test infrastructure only).  Anchor generation restated from the reference.

`generate_anchors` follows generate_anchors.py:42-109 (importable; pinned against it and
against its known-answer comment generate_anchors.py:30-38 by tests/test_oracle_anchors.py);
`get_all_anchors` follows utils.py:606-658 (utils.py itself imports TensorFlow at module
top, so it cannot be imported here -- restated).
"""
from __future__ import annotations

import numpy as np


def _whctrs(anchor):                                   # generate_anchors.py:60-69
    w = anchor[2] - anchor[0] + 1
    h = anchor[3] - anchor[1] + 1
    return w, h, anchor[0] + 0.5 * (w - 1), anchor[1] + 0.5 * (h - 1)


def _mkanchors(ws, hs, x_ctr, y_ctr):                  # generate_anchors.py:71-83
    ws = ws[:, np.newaxis]
    hs = hs[:, np.newaxis]
    return np.hstack((x_ctr - 0.5 * (ws - 1), y_ctr - 0.5 * (hs - 1),
                      x_ctr + 0.5 * (ws - 1), y_ctr + 0.5 * (hs - 1)))


def generate_anchors(base_size=16, ratios=(0.5, 1, 2), scales=2 ** np.arange(3, 6)):
    """generate_anchors.py:42-58: ratio enumeration then scale enumeration around a
    (0,0,base-1,base-1) window; widths/heights rounded with np.round."""
    ratios = np.asarray(ratios, dtype=np.float64)
    scales = np.asarray(scales, dtype=np.float64)
    base = np.array([1, 1, base_size, base_size], dtype="float32") - 1
    w, h, xc, yc = _whctrs(base)
    size_ratios = (w * h) / ratios                     # :90-96
    ws = np.round(np.sqrt(size_ratios))
    hs = np.round(ws * ratios)
    ratio_anchors = _mkanchors(ws, hs, xc, yc)
    out = []
    for i in range(ratio_anchors.shape[0]):            # :101-109
        w, h, xc, yc = _whctrs(ratio_anchors[i])
        out.append(_mkanchors(w * scales, h * scales, xc, yc))
    return np.vstack(out)


def get_all_anchors(stride, sizes, ratios, max_size):
    """utils.py:606-658: field of anchors [S,S,A,4] float32 for S = ceil(max_size/stride),
    shifts k*stride, and +1 on x2,y2 (utils.py:657)."""
    cell = generate_anchors(stride, ratios=np.asarray(ratios, dtype=np.float64),
                            scales=np.asarray(sizes, dtype=np.float64) / stride)
    field = int(np.ceil(max_size / stride))
    shifts = np.arange(0, field) * stride
    sx, sy = np.meshgrid(shifts, shifts)
    sx = sx.flatten(); sy = sy.flatten()
    shifts = np.vstack((sx, sy, sx, sy)).transpose()
    K = shifts.shape[0]
    A = cell.shape[0]
    f = cell.reshape((1, A, 4)) + shifts.reshape((1, K, 4)).transpose((1, 0, 2))
    f = f.reshape((field, field, A, 4)).astype("float32")
    f[:, :, :, [2, 3]] += 1
    return f
