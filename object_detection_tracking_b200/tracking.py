"""GPU drop-in for the DeepSORT appearance metric.

`GpuNearestNeighborDistanceMetric` has the constructor, attributes and methods of the reference's
`deep_sort.nn_matching.NearestNeighborDistanceMetric` (deep_sort/nn_matching.py:99-177), so
`Tracker(metric)` (deep_sort/tracker.py:40-48, obj_detect_tracking.py:553-558) takes it unchanged.
`distance()` replaces the per-track Python loop of tiny NumPy GEMMs (:174-177) by ONE tensor-core GEMM
over the concatenated galleries plus a segmented row-min (b2_cosine_cost).  No CPU fallback.
"""
from __future__ import annotations

import numpy as np

from .engine import cosine_cost


class GpuNearestNeighborDistanceMetric(object):
    def __init__(self, metric, matching_threshold, budget=None, device=0, precision="split"):
        if metric != "cosine":
            raise ValueError("Invalid metric; the B200 path implements 'cosine' (the metric the drivers use, "
                             "obj_detect_tracking.py:553)")
        self.matching_threshold = matching_threshold
        self.budget = budget
        self.samples = {}
        self.device = device
        self.precision = precision

    def partial_fit(self, features, targets, active_targets):
        """deep_sort/nn_matching.py:137-154 (host bookkeeping, unchanged semantics)."""
        for feature, target in zip(features, targets):
            self.samples.setdefault(target, []).append(feature)
            if self.budget is not None:
                self.samples[target] = self.samples[target][-self.budget:]
        self.samples = {k: self.samples[k] for k in active_targets}

    def distance(self, features, targets):
        """Returns the float64 [len(targets), len(features)] cost matrix of :156-177."""
        T, N = len(targets), len(features)
        cost = np.zeros((T, N))
        if T == 0 or N == 0:
            return cost
        feats = np.asarray(features, dtype=np.float32)
        seg = np.zeros(T + 1, dtype=np.int32)
        rows = []
        for i, t in enumerate(targets):
            g = np.asarray(self.samples[t], dtype=np.float32).reshape(-1, feats.shape[1])
            rows.append(g)
            seg[i + 1] = seg[i] + g.shape[0]
        gallery = np.concatenate(rows, axis=0)
        cost[:, :] = cosine_cost(gallery, seg, feats, device=self.device, precision=self.precision)
        return cost
