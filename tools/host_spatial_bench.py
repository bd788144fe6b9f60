"""Host-only timing of the multi-camera trajectory distance (multi_video_reid.py:260-305): native b2_track_spatial_dist
against the reference's compute_spatial_dist on the same tracks.  Needs /root/reference.  One JSON line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def main():
    import make_golden_tmot as mg
    mg.install_stubs()
    from unittest.mock import MagicMock
    for m in ("utils", "enqueuer_thread", "diva_io", "diva_io.video", "moviepy", "moviepy.editor", "cv2", "tqdm",
              "torchreid", "torchreid.feature_extractor", "torchreid.distance"):
        sys.modules.setdefault(m, MagicMock())
    import multi_video_reid as mvr
    from object_detection_tracking_b200 import reid
    rng = np.random.default_rng(3)

    def camera(n, first):
        out = {}
        for k in range(n):
            f0, ln = int(rng.integers(0, 200)), int(rng.integers(50, 300))
            rows = np.zeros((ln, 9))
            rows[:, 0] = np.arange(f0, f0 + ln)
            rows[:, -2:] = rng.uniform(0, 300, 2) + rng.uniform(-1, 1, 2) * np.arange(ln)[:, None]
            out[first + k] = (rows, np.zeros((1, 8), np.float32))
        return out
    c1, c2 = camera(50, 1), camera(50, 1000)
    t0 = time.perf_counter()
    ref = mvr.compute_spatial_dist(c1, c2, frame_offset=3, tol=50)
    t_ref = time.perf_counter() - t0
    reid.compute_spatial_dist(c1, c2, 3, 50)
    t0 = time.perf_counter()
    got = reid.compute_spatial_dist(c1, c2, 3, 50)
    t_nat = time.perf_counter() - t0
    print(json.dumps(dict(what="spatial_dist_50x50_tracks_host_only", reference_ms=round(t_ref * 1e3, 2),
                          native_ms=round(t_nat * 1e3, 2), speedup=round(t_ref / t_nat, 1),
                          same_gate=bool(np.array_equal(got < 9999, ref < 9999)), max_abs_diff=float(np.abs(got - ref).max()),
                          comparable_pairs=int((ref < 9999).sum()))))


if __name__ == "__main__":
    main()
