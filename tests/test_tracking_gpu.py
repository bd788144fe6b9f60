"""DeepSORT appearance cost on the GPU vs golden vectors produced by the reference's own nn_matching."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_cosine_cost_matches_reference_golden(golden_dir):
    from object_detection_tracking_b200.engine import cosine_cost
    g = np.load(os.path.join(golden_dir, "deepsort_cosine.npz"))
    cost = cosine_cost(g["gallery"], g["seg"], g["dets"], precision="split")
    assert cost.shape == g["cost"].shape
    assert np.abs(cost - g["cost"]).max() < 2e-6       # fp32 cosine in [0, 2]: a few ulps
    c16 = cosine_cost(g["gallery"], g["seg"], g["dets"], precision="fp16")
    assert np.abs(c16 - g["cost"]).max() < 2e-3


def test_metric_class_is_drop_in(golden_dir):
    from object_detection_tracking_b200.tracking import GpuNearestNeighborDistanceMetric
    from oracle import nn_matching
    rng = np.random.default_rng(0)
    m_gpu = GpuNearestNeighborDistanceMetric("cosine", 0.5, budget=5)
    m_ref = nn_matching.NearestNeighborDistanceMetric("cosine", 0.5, budget=5)
    for step in range(4):
        feats = np.abs(rng.standard_normal((9, 256))).astype(np.float32)
        targets = rng.integers(0, 6, 9)
        active = sorted(set(targets.tolist()))
        m_gpu.partial_fit(feats, targets, active)
        m_ref.partial_fit(feats, targets, active)
        q = np.abs(rng.standard_normal((11, 256))).astype(np.float32)
        a = m_gpu.distance(q, active)
        b = m_ref.distance(q, active)
        assert a.dtype == np.float64 and a.shape == b.shape
        assert np.abs(a - b).max() < 2e-6
    assert m_gpu.distance(np.zeros((0, 256), np.float32), [0]).shape == (1, 0)     # empty detections
    with pytest.raises(ValueError):
        GpuNearestNeighborDistanceMetric("manhattan", 0.5)                           # nn_matching.py:123-129
    with pytest.raises(ValueError):
        nn_matching.NearestNeighborDistanceMetric("manhattan", 0.5)


def test_euclidean_metric_matches_reference_golden(golden_dir):
    from object_detection_tracking_b200.tracking import GpuNearestNeighborDistanceMetric
    g = np.load(os.path.join(golden_dir, "deepsort_euclid.npz"))
    m = GpuNearestNeighborDistanceMetric("euclidean", 0.3, int(g["budget"]))
    m.partial_fit(g["feats"], g["targets"], list(range(int(g["T"]))))
    cost = m.distance(g["dets"], g["order"].tolist())
    assert np.abs(cost - g["cost"]).max() <= 5e-6 * np.abs(g["cost"]).max()


def test_euclidean_metric_matches_oracle():
    """NearestNeighborDistanceMetric('euclidean') (deep_sort/nn_matching.py:5-28,57-75,156-177): min over a track's
    gallery rows of the squared euclidean distance, clamped at 0 -- against the oracle class on the same samples."""
    from object_detection_tracking_b200.tracking import GpuNearestNeighborDistanceMetric
    from oracle import nn_matching
    rng = np.random.default_rng(14)
    m_gpu = GpuNearestNeighborDistanceMetric("euclidean", 0.3, budget=4)
    m_ref = nn_matching.NearestNeighborDistanceMetric("euclidean", 0.3, budget=4)
    for step in range(3):
        feats = rng.standard_normal((30, 128)).astype(np.float32)
        targets = [i % 6 for i in range(30)]
        m_gpu.partial_fit(feats, targets, list(range(6)))
        m_ref.partial_fit(feats, targets, list(range(6)))
        q = rng.standard_normal((11, 128)).astype(np.float32)
        a = m_gpu.distance(q, [4, 0, 2])
        b = m_ref.distance(q, [4, 0, 2])
        assert a.dtype == np.float64 and a.shape == b.shape == (3, 11)
        assert np.abs(a - b).max() <= 5e-6 * np.abs(b).max()


def test_large_gallery_many_tiles():
    """budget 60 x 100 tracks (multi-queuer defaults): S = 6000 rows, 47 M-tiles."""
    from object_detection_tracking_b200.engine import cosine_cost
    from oracle import nn_matching
    rng = np.random.default_rng(1)
    T, S_per, N, D = 100, 60, 100, 256
    gal = np.abs(rng.standard_normal((T * S_per, D))).astype(np.float32)
    dets = np.abs(rng.standard_normal((N, D))).astype(np.float32)
    seg = np.arange(0, T * S_per + 1, S_per, dtype=np.int32)
    cost = cosine_cost(gal, seg, dets)
    ref = np.stack([nn_matching.nn_cosine_distance(gal[seg[t]:seg[t + 1]], dets) for t in range(T)])
    assert np.abs(cost - ref).max() < 2e-6


def test_track_ids_with_gpu_metric_equal_reference_run(golden_dir):
    """The whole association loop with the GPU appearance metric reproduces the reference tracker's
    ids bit-exactly and its boxes to 1e-6 on the golden sequence (ids depend on thresholded costs)."""
    from object_detection_tracking_b200.tracking import GpuNearestNeighborDistanceMetric
    from oracle import deepsort
    g = np.load(os.path.join(golden_dir, "deepsort_tracker.npz"))
    n = len([k for k in g.files if k.startswith("frame")])
    frames = [g["frame%d" % i] for i in range(n)]
    got = deepsort.run_sequence(frames, GpuNearestNeighborDistanceMetric("cosine", 0.5, 5))
    assert got.shape == g["results"].shape
    np.testing.assert_array_equal(got[:, :2], g["results"][:, :2])
    assert np.abs(got[:, 2:] - g["results"][:, 2:]).max() < 1e-6


def test_native_tracker_with_gpu_appearance_cost_equals_reference_run(golden_dir):
    """csrc/tracker.cpp end to end (Kalman, cascade, assignment in native code; appearance cost = b2_cosine_cost on the GPU,
    the default): ids and life cycle of the reference's own Tracker on both golden sequences."""
    from object_detection_tracking_b200.tracking import GpuNearestNeighborDistanceMetric, Tracker
    from oracle import deepsort
    for name in ("deepsort_tracker.npz", "deepsort_tracker_crowd.npz"):
        g = np.load(os.path.join(golden_dir, name))
        n = len([k for k in g.files if k.startswith("frame")])
        trk = Tracker(GpuNearestNeighborDistanceMetric("cosine", 0.5, 5))
        rows_out = []
        for f in range(n):
            dets = [deepsort.Detection(r[:4], r[4], r[5:]) for r in g["frame%d" % f]]
            trk.predict()
            trk.update(dets)
            for t in trk.tracks:
                if t.is_confirmed() and t.time_since_update <= 1:
                    rows_out.append([f, t.track_id] + t.to_tlwh().tolist())
        got = np.asarray(rows_out, dtype=np.float64).reshape(-1, 6)
        assert got.shape == g["results"].shape
        np.testing.assert_array_equal(got[:, :2], g["results"][:, :2])
        assert np.abs(got[:, 2:] - g["results"][:, 2:]).max() < 1e-6
        trk.close()


def test_native_tracker_honours_the_euclidean_metric(golden_dir):
    """Tracker(NearestNeighborDistanceMetric('euclidean', thr)): the reference Tracker asks metric.distance()
    (deep_sort/tracker.py:98-104), so the native cascade must use the squared-euclidean cost too.  Checked
    against the oracle's association loop driven by the oracle's euclidean metric on the crowd sequence (features
    rescaled to unit norm so that a squared-distance threshold of 0.6 is meaningful)."""
    from object_detection_tracking_b200.tracking import GpuNearestNeighborDistanceMetric, Tracker
    from oracle import deepsort, nn_matching
    g = np.load(os.path.join(golden_dir, "deepsort_tracker_crowd.npz"))
    n = len([k for k in g.files if k.startswith("frame")])
    frames = []
    for f in range(n):
        r = g["frame%d" % f].copy()
        r[:, 5:] /= np.linalg.norm(r[:, 5:], axis=1, keepdims=True)
        frames.append(r)
    ref = deepsort.run_sequence(frames, nn_matching.NearestNeighborDistanceMetric("euclidean", 0.6, 5))
    ref_cos = deepsort.run_sequence(frames, nn_matching.NearestNeighborDistanceMetric("cosine", 0.6, 5))
    trk = Tracker(GpuNearestNeighborDistanceMetric("euclidean", 0.6, 5))
    rows_out = []
    for f, rows in enumerate(frames):
        trk.predict()
        trk.update([deepsort.Detection(r[:4], r[4], r[5:]) for r in rows])
        for t in trk.tracks:
            if t.is_confirmed() and t.time_since_update <= 1:
                rows_out.append([f, t.track_id] + t.to_tlwh().tolist())
    trk.close()
    got = np.asarray(rows_out, dtype=np.float64).reshape(-1, 6)
    assert got.shape == ref.shape
    np.testing.assert_array_equal(got[:, :2], ref[:, :2])
    assert np.abs(got[:, 2:] - ref[:, 2:]).max() < 1e-6
    # the two metrics really associate differently on this sequence, so the test can tell them apart
    assert ref.shape != ref_cos.shape or not np.array_equal(ref[:, :2], ref_cos[:, :2])
