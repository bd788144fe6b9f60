"""The DeepSORT appearance-cost path (normalise -> one tcgen05 GEMM -> segmented min) at the reference's working size
(100 tracks x budget 5, 100 detections, D = 256), for an ncu capture.  A number printed under ncu is never a bench value."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from object_detection_tracking_b200.engine import cosine_cost  # noqa: E402

rng = np.random.default_rng(0)
T, S, N, D = 100, 5, 100, 256
gal = np.abs(rng.standard_normal((T * S, D))).astype(np.float32)
dets = np.abs(rng.standard_normal((N, D))).astype(np.float32)
seg = (np.arange(T + 1) * S).astype(np.int32)
for _ in range(2):
    cost = cosine_cost(gal, seg, dets)
print("cost", cost.shape, float(cost.min()), float(cost.max()))
