"""Empty / degenerate inputs of the host-side entry points (no GPU needed: they return before any device call)."""
import numpy as np

from object_detection_tracking_b200 import _lib, reid, tmot, tracking


def test_lapjv_empty_and_single():
    opt, x, y = tmot.lapjv(np.zeros((0, 3)), cost_limit=1.0)
    assert opt == 0 and len(x) == 0 and list(y) == [-1, -1, -1]
    opt, x, y = tmot.lapjv(np.array([[0.2]]), cost_limit=1.0)
    assert list(x) == [0] and list(y) == [0] and abs(opt - 0.2) < 1e-15
    opt, x, y = tmot.lapjv(np.array([[5.0]]), cost_limit=1.0)          # above the limit: stays unmatched
    assert list(x) == [-1] and list(y) == [-1] and opt == 0
    opt, x, y = tmot.lapjv(np.full((3, 2), np.inf), cost_limit=0.7)
    assert list(x) == [-1, -1, -1] and list(y) == [-1, -1]


def test_distances_with_no_rows():
    assert tmot.iou_distance([], []).shape == (0, 0)
    assert tmot.iou_distance([np.array([0., 0., 5., 5.])], []).shape == (1, 0)
    assert tmot.fuse_motion(np.zeros((0, 8)), np.zeros((0, 8, 8)), np.zeros((0, 4)), np.zeros((4, 4))).shape == (0, 4)
    assert tmot.embedding_distance(np.zeros((0, 16), np.float32), np.zeros((3, 16), np.float32)).shape == (0, 3)
    d = tmot.iou_distance([np.array([0., 0., 9., 9.])], [np.array([0., 0., 9., 9.]), np.array([100., 100., 110., 110.])])
    assert d[0, 0] == 0.0 and d[0, 1] == 1.0


def test_pair_cost_and_spatial_degenerate():
    lib = _lib.load()
    out = np.zeros((0, 0), np.float32)
    seg = np.zeros(1, np.int32)
    assert lib.b2_track_pair_cost(0, None, _lib.ptr(seg), 0, None, _lib.ptr(seg), 0, 8, None, 999.0, 1, _lib.ptr(out)) == 0
    # tracks without a single crop on either side: every pair is `fill`, decided on the host
    seg_a = np.array([0, 0, 0], np.int32)
    seg_b = np.array([0, 0], np.int32)
    a = np.zeros((1, 8), np.float32)
    out = np.zeros((2, 1), np.float32)
    assert lib.b2_track_pair_cost(0, _lib.ptr(a), _lib.ptr(seg_a), 2, _lib.ptr(a), _lib.ptr(seg_b), 1, 8, None, 999.0, 1,
                                  _lib.ptr(out)) == 0
    assert (out == 999.0).all()
    bad = np.array([0, 3, 2], np.int32)                                  # not monotone: refused, with a message
    assert lib.b2_track_pair_cost(0, _lib.ptr(a), _lib.ptr(bad), 2, _lib.ptr(a), _lib.ptr(seg_b), 1, 8, None, 999.0, 1,
                                  _lib.ptr(out)) != 0
    assert b"monotone" in lib.b2_last_error()
    t1 = {1: (np.array([[0, 0, 0, 0, 0, 0, 0, 5., 5.]]), np.zeros((1, 8), np.float32))}
    t2 = {7: (np.array([[3, 0, 0, 0, 0, 0, 0, 5., 5.]]), np.zeros((1, 8), np.float32))}
    assert (reid.compute_spatial_dist(t1, t2) == 9999.).all()            # no common frame
    assert reid.compute_spatial_dist(t1, t2, frame_offset=-3)[0, 0] == 0.0
    assert reid.match_tracks({}, t2) == [] and reid.match_tracks(t1, {}) == []


def test_glue_with_no_detections():
    z4, z = np.zeros((0, 4), np.float32), np.zeros((0,), np.float32)
    assert tracking.create_obj_infos(0, z4, z, np.zeros((0,), np.int64), np.zeros((0, 256, 7, 7), np.float32), {}, ["Person"],
                                     0.5, 0, 1.0) == []
    assert tracking.preprocess_detections(z4, z, np.zeros((0,), np.int64), np.zeros((0, 256), np.float32), {}, ["Person"],
                                          0.5, 1.0) == []
    assert tracking.non_max_suppression(np.zeros((0, 4)), 0.85, np.zeros(0)) == []


def test_jde_frames_without_detections_before_the_first_one():
    """frame_id keeps counting through empty frames (multitracker.py:228), so a track's start_frame is the real frame."""
    trk = tmot.JDETracker(0.5, id_group=tmot._IdGroup(), cost_fn=lambda a, b: np.zeros((len(a), len(b))))
    assert trk.update([]) == [] and trk.update([]) == []
    det = [(np.array([10., 10., 40., 80.]), 0.9, np.ones(8, np.float32))]
    trk.update(det)
    out = trk.update(det)
    assert len(out) == 1 and out[0].start_frame == 3 and out[0].frame_id == 4 and out[0].track_id == 1
    trk.close()


def test_jde_id_counter_survives_the_first_tracker_of_a_group():
    """BaseTrack._count outlives any single JDETracker: a tracker that joins after the first one was closed still draws
    from the same counter as the remaining members."""
    grp = tmot._IdGroup()
    cost = lambda a, b: np.zeros((len(a), len(b)))
    det = lambda x: [(np.array([x, 10., 40., 80.]), 0.9, np.ones(8, np.float32))]
    a = tmot.JDETracker(0.5, id_group=grp, cost_fn=cost)
    b = tmot.JDETracker(0.5, id_group=grp, cost_fn=cost)
    a.update(det(10.))            # id 1
    b.update(det(500.))           # id 2
    a.close()
    c = tmot.JDETracker(0.5, id_group=grp, cost_fn=cost)
    c.update(det(900.))           # id 3: joined through b
    assert [t.track_id for t in c.get_tracks(1)] == [3] and [t.track_id for t in b.get_tracks(1)] == [2]
    b.close()
    c.close()
