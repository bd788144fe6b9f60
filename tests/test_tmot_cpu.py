"""Native TMOT / JDE association (csrc/tmot.cpp, host code) against fixtures produced by the REFERENCE's own
tmot/multitracker.py + tmot/matching.py + tmot/kalman_filter.py (tests/golden/make_golden_tmot.py; lap / cython_bbox /
numba.jit stubbed there as documented).  The embedding distance is supplied by a float64 numpy checker here; the product
path computes it with b2_distance_matrix on the GPU (tests/test_widen_gpu.py)."""
import os

import numpy as np
import pytest
import scipy.optimize


def cdist_cost(a, b):          # scipy.spatial.distance.cdist(a, b) restated (matching.py:92)
    a, b = a.astype(np.float64), b.astype(np.float64)
    return np.sqrt(np.maximum(0.0, ((a[:, None, :] - b[None, :, :]) ** 2).sum(-1)))


def test_jde_tracker_reproduces_reference_run(golden_dir):
    """80 frames x 2 trackers sharing the id counter: embedding + Kalman-gated association, IoU rounds, unconfirmed
    tracks, low-confidence detections, lost / re-found / removed tracks, duplicate removal."""
    from object_detection_tracking_b200.tmot import JDETracker, _IdGroup
    g = np.load(os.path.join(golden_dir, "tmot_jde.npz"))
    grp = _IdGroup()
    trackers = [JDETracker(0.5, track_max_second_lost=4.0, frame_gap=8., frame_rate=30., cost_fn=cdist_cost, id_group=grp)
                for _ in range(2)]
    out_rows, list_rows = [], []
    for f in range(80):
        for k, trk in enumerate(trackers):
            fr = g["s%d_f%d" % (k, f)]
            outs = trk.update([(r[:4].astype(np.float64), float(r[4]), r[5:].copy()) for r in fr])
            for t in outs:
                out_rows.append([f, k, t.track_id] + t.tlwh.tolist() + t.cur_det_tlwh.tolist()
                                + [t.cur_det_conf, t.score, t.tracklet_len, t.start_frame])
            for which in (1, 2):
                for t in trk.get_tracks(which):
                    list_rows.append([f, k, which, t.track_id, t.state, int(t.is_activated), t.frame_id])
    got = np.asarray(out_rows, dtype=np.float64)
    lists = np.asarray(list_rows, dtype=np.int64)
    assert got.shape == g["out"].shape
    np.testing.assert_array_equal(got[:, :3], g["out"][:, :3])                 # frame, tracker, track id: bit-exact
    np.testing.assert_array_equal(got[:, 7:], g["out"][:, 7:])                 # detection box / conf / score / lengths
    assert np.abs(got[:, 3:7] - g["out"][:, 3:7]).max() < 1e-8                 # Kalman-filtered box (float64)
    np.testing.assert_array_equal(lists, g["lists"])                           # tracked / lost lists: ids, states, order
    assert (lists[:, 2] == 2).any() and (lists[:, 4] == 3).any()               # lost list and the Removed-state quirk seen
    fin = trackers[0].get_tracks(1)
    np.testing.assert_array_equal([t.track_id for t in fin], g["final_ids"])
    assert np.abs(np.asarray([t.mean for t in fin]) - g["final_mean"]).max() < 1e-8
    assert np.abs(np.asarray([t.covariance for t in fin]) - g["final_cov"]).max() < 1e-8
    trackers[0].reset()
    assert trackers[0].get_tracks(1) == [] and trackers[0].update([]) == []
    for t in trackers:
        t.close()


def test_reset_of_a_tracker_without_state_zeroes_the_shared_id_counter(golden_dir):
    """multitracker.py:206-215: reset() of ANY JDETracker sets BaseTrack._count = 0 -- also of one that has not seen a
    detection yet (the vehicle tracker of a video without vehicles); ids of the next video then restart at 1."""
    from object_detection_tracking_b200.tmot import JDETracker, _IdGroup
    g = np.load(os.path.join(golden_dir, "tmot_jde.npz"))
    grp = _IdGroup()
    person, vehicle = (JDETracker(0.5, cost_fn=cdist_cost, id_group=grp) for _ in range(2))
    frames = [[(r[:4].astype(np.float64), float(r[4]), r[5:].copy()) for r in g["s0_f%d" % f]] for f in range(3)]
    first = [t.track_id for f in frames for t in person.update(f)]
    assert first and min(first) == 1
    vehicle.reset()                                       # owns no native handle yet
    person2 = JDETracker(0.5, cost_fn=cdist_cost, id_group=grp)
    again = [t.track_id for f in frames for t in person2.update(f)]
    assert again == first                                 # counter restarted: the same ids as the first video
    for t in (person, vehicle, person2):
        t.close()


def test_matching_functions_match_reference(golden_dir):
    from object_detection_tracking_b200 import tmot
    g = np.load(os.path.join(golden_dir, "tmot_matching.npz"))
    np.testing.assert_allclose(cdist_cost(g["track_feats"], g["det_feats"]), g["emb"], rtol=0, atol=1e-12)
    for only_pos in (0, 1):
        fused = tmot.fuse_motion(g["means"], g["covs"], g["emb"], g["xyah"], only_position=bool(only_pos))
        ref = g["fused_%d" % only_pos]
        np.testing.assert_array_equal(np.isinf(fused), np.isinf(ref))
        fin = np.isfinite(ref)
        assert fin.any() and np.isinf(ref).any()
        assert np.abs(fused[fin] - ref[fin]).max() < 1e-9
    np.testing.assert_allclose(tmot.iou_distance(list(g["tlbr_a"]), list(g["tlbr_b"])), g["iou_dist"], rtol=0, atol=1e-15)
    for k in range(3):
        cost = g["fused_0"] if k == 0 else g["iou_dist"]
        m, ua, ub = tmot.linear_assignment(cost, float(g["la%d_thr" % k]))
        np.testing.assert_array_equal(np.asarray(m).reshape(-1, 2), g["la%d_matches" % k])
        np.testing.assert_array_equal(ua, g["la%d_ua" % k])
        np.testing.assert_array_equal(ub, g["la%d_ub" % k])
    assert len(g["la0_matches"]) > 0 and len(g["la2_ua"]) > 0
    m, ua, ub = tmot.linear_assignment(np.zeros((0, 4)), 0.5)
    assert m.shape == (0, 2) and tuple(ua) == () and tuple(ub) == (0, 1, 2, 3)


def test_lapjv_cost_limit_is_the_optimum_of_the_extended_problem():
    """lap.lapjv(extend_cost=True, cost_limit=L): random rectangular matrices (tie-free), including inf entries and the
    999-filled gated matrices of multi_video_reid.py:308-324,512 -- matched set = optimum of the (nr+nc)^2 extension."""
    from object_detection_tracking_b200 import tmot
    rng = np.random.default_rng(11)
    for case in range(300):
        nr, nc = int(rng.integers(1, 14)), int(rng.integers(1, 14))
        cost = rng.uniform(0, 2, (nr, nc))
        limit = float(rng.uniform(0.3, 1.6))
        if case % 3 == 0:
            cost[rng.uniform(size=cost.shape) < 0.3] = np.inf
        if case % 3 == 1:
            cost = np.where(rng.uniform(size=cost.shape) < 0.5, 999.0, cost * 300)
            limit = 998.0
        opt, x, y = tmot.lapjv(cost, extend_cost=True, cost_limit=limit)
        n = nr + nc
        ext = np.full((n, n), limit / 2)
        ext[nr:, nc:] = 0
        ext[:nr, :nc] = cost
        r, c = scipy.optimize.linear_sum_assignment(ext)
        ref_pairs = sorted((int(i), int(j)) for i, j in zip(r, c) if i < nr and j < nc)
        got_pairs = sorted((i, int(j)) for i, j in enumerate(x) if j >= 0)
        assert got_pairs == ref_pairs
        for i, j in got_pairs:
            assert y[j] == i and cost[i, j] < limit
        assert sorted(np.where(y < 0)[0]) == sorted(set(range(nc)) - set(j for _, j in got_pairs))
        assert abs(opt - sum(cost[i, j] for i, j in got_pairs)) < 1e-9


def test_jde_needs_gpu_without_cost_fn():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from object_detection_tracking_b200.tmot import JDETracker, _IdGroup
    rng = np.random.default_rng(0)
    trk = JDETracker(0.5, id_group=_IdGroup())
    dets = [(np.array([10., 10., 40., 80.]), 0.9, rng.standard_normal(16).astype(np.float32))]
    trk.update(dets)                      # first frame: no tracks yet -> no embedding distance needed
    trk.update(dets)                      # second: unconfirmed track, IoU only
    with pytest.raises(RuntimeError):
        trk.update(dets)                  # third: activated track -> embedding distance -> needs the GPU
    trk.close()


def test_jde_tracker_other_parameters(golden_dir):
    """Second reference run: short lost window (7.5 frames), tighter gates, alpha 0.6, confidence gate 0.6."""
    from object_detection_tracking_b200.tmot import JDETracker, _IdGroup
    g = np.load(os.path.join(golden_dir, "tmot_jde_b.npz"))
    trk = JDETracker(0.6, track_max_second_lost=1.0, emb_max_dist=0.5, iou_max_dist1=0.6, iou_max_dist2=0.7,
                     emb_smooth_alpha=0.6, frame_gap=4., frame_rate=30., cost_fn=cdist_cost, id_group=_IdGroup())
    out_rows, list_rows = [], []
    for f in range(60):
        fr = g["f%d" % f]
        for t in trk.update([(r[:4].astype(np.float64), float(r[4]), r[5:].copy()) for r in fr]):
            out_rows.append([f, t.track_id] + t.tlwh.tolist() + [t.score, t.tracklet_len, t.start_frame])
        for which in (1, 2):
            for t in trk.get_tracks(which):
                list_rows.append([f, which, t.track_id, t.state, int(t.is_activated), t.frame_id])
    got = np.asarray(out_rows, dtype=np.float64)
    assert got.shape == g["out"].shape
    np.testing.assert_array_equal(got[:, :2], g["out"][:, :2])
    np.testing.assert_array_equal(got[:, 6:], g["out"][:, 6:])
    assert np.abs(got[:, 2:6] - g["out"][:, 2:6]).max() < 1e-8
    np.testing.assert_array_equal(np.asarray(list_rows, dtype=np.int64), g["lists"])
    trk.close()
