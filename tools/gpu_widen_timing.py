"""Host-clock timings of the 8f widening on a GPU box (not bench values: small synchronous calls, host clock):
JDE tracker per frame (native + GPU embedding distance) vs the same loop with a numpy cost, the multi-camera pair cost
vs a per-pair numpy loop, and the device-side frame resize vs cv2 on the host.  One JSON line each."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def jde():
    from object_detection_tracking_b200.tmot import JDETracker, _IdGroup
    rng = np.random.default_rng(0)
    n_obj, D, frames = 100, 256, 60
    proto = rng.standard_normal((n_obj, D)).astype(np.float32)
    pos = rng.uniform(50, 1800, (n_obj, 2))
    vel = rng.uniform(-6, 6, (n_obj, 2))

    def dets(f):
        return [(np.concatenate([pos[o] + vel[o] * f, [40., 90.]]), 0.9,
                 proto[o] + 0.1 * rng.standard_normal(D).astype(np.float32)) for o in range(n_obj)]

    def cdist(a, b):
        a, b = a.astype(np.float64), b.astype(np.float64)
        return np.sqrt(np.maximum(0, (a * a).sum(1)[:, None] + (b * b).sum(1)[None] - 2 * a @ b.T))
    res = {}
    for name, kw in (("gpu_cost", {}), ("numpy_cost", {"cost_fn": cdist})):
        trk = JDETracker(0.5, id_group=_IdGroup(), **kw)
        ids = []
        for f in range(5):
            trk.update(dets(f))
        t0 = time.perf_counter()
        for f in range(5, 5 + frames):
            ids.append(sorted(t.track_id for t in trk.update(dets(f))))
        res[name + "_ms_per_frame"] = (time.perf_counter() - t0) / frames * 1e3
        res[name + "_ids"] = ids
        trk.close()
    same = res.pop("gpu_cost_ids") == res.pop("numpy_cost_ids")
    print(json.dumps(dict(what="jde_tracker_100obj_D256", same_ids=same, **res)))


def pair_cost():
    from object_detection_tracking_b200 import _lib
    rng = np.random.default_rng(1)
    N, M, D = 50, 50, 512
    ka, kb = rng.integers(20, 100, N), rng.integers(20, 100, M)
    sa = np.concatenate([[0], np.cumsum(ka)]).astype(np.int32)
    sb = np.concatenate([[0], np.cumsum(kb)]).astype(np.int32)
    a = rng.standard_normal((sa[-1], D)).astype(np.float32)
    b = rng.standard_normal((sb[-1], D)).astype(np.float32)
    out = np.zeros((N, M), np.float32)
    lib = _lib.load()
    for _ in range(2):
        _lib.check(lib.b2_track_pair_cost(0, _lib.ptr(a), _lib.ptr(sa), N, _lib.ptr(b), _lib.ptr(sb), M, D, None, 999.0, 1, _lib.ptr(out)))
    t0 = time.perf_counter()
    for _ in range(5):
        _lib.check(lib.b2_track_pair_cost(0, _lib.ptr(a), _lib.ptr(sa), N, _lib.ptr(b), _lib.ptr(sb), M, D, None, 999.0, 1, _lib.ptr(out)))
    gpu_ms = (time.perf_counter() - t0) / 5 * 1e3
    from sklearn.metrics.pairwise import euclidean_distances
    t0 = time.perf_counter()
    ref = np.zeros((N, M))
    for i in range(N):
        for j in range(M):
            ref[i, j] = euclidean_distances(a[sa[i]:sa[i + 1]], b[sb[j]:sb[j + 1]], squared=True).min()
    cpu_ms = (time.perf_counter() - t0) * 1e3
    print(json.dumps(dict(what="pair_cost_50x50_tracks", crops=[int(sa[-1]), int(sb[-1])], gpu_ms=gpu_ms, sklearn_loop_ms=cpu_ms,
                          max_rel_err=float(np.abs(out - ref).max() / np.abs(ref).max()))))


def resize():
    import cv2
    from object_detection_tracking_b200.engine import resize_frames
    rng = np.random.default_rng(2)
    src = rng.integers(0, 256, (8, 1080, 1920, 3)).astype(np.uint8)
    resize_frames(src, 1280, 720)
    t0 = time.perf_counter()
    for _ in range(5):
        resize_frames(src, 1280, 720)
    gpu_ms = (time.perf_counter() - t0) / 5 * 1e3
    t0 = time.perf_counter()
    for f in src:
        cv2.resize(f.astype("float32"), (1280, 720), interpolation=cv2.INTER_LINEAR)
    cpu_ms = (time.perf_counter() - t0) * 1e3
    print(json.dumps(dict(what="resize_8x1080p_to_720p_incl_copies", gpu_ms=gpu_ms, cv2_host_ms=cpu_ms)))


if __name__ == "__main__":
    for fn in (jde, pair_cost, resize):
        try:
            fn()
        except Exception as e:          # keep going: each line is independent
            print(json.dumps(dict(what=fn.__name__, error=repr(e))))
