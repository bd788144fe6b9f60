"""Native DeepSORT association loop (csrc/tracker.cpp, host code) against fixtures produced by the REFERENCE's own
deep_sort.Tracker / preprocessing.non_max_suppression (tests/golden/make_golden.py) and against SciPy's
linear_sum_assignment.  The appearance cost is supplied by the CPU oracle here (checker role); the product path uses
b2_cosine_cost on the GPU (tests/test_tracking_gpu.py)."""
import os

import numpy as np
import pytest

from oracle import deepsort, nn_matching


def oracle_cost(gallery, seg, dets):
    return np.stack([nn_matching.nn_cosine_distance(gallery[seg[t]:seg[t + 1]], dets) for t in range(len(seg) - 1)])


def run_native(frames, **kw):
    from object_detection_tracking_b200.tracking import Tracker
    metric = nn_matching.NearestNeighborDistanceMetric("cosine", 0.5, 5)       # parameters only
    trk = Tracker(metric, cost_fn=oracle_cost, **kw)
    results, life = [], []
    for f, rows in enumerate(frames):
        dets = [deepsort.Detection(r[:4], r[4], r[5:]) for r in rows]
        trk.predict()
        trk.update(dets)
        for t in trk.tracks:
            life.append([f, t.track_id, t.state, t.hits, t.age, t.time_since_update])
            if t.is_confirmed() and t.time_since_update <= 1:
                results.append([f, t.track_id] + t.to_tlwh().tolist())
    trk.close()
    return np.asarray(results, dtype=np.float64).reshape(-1, 6), np.asarray(life, dtype=np.int64).reshape(-1, 6)


def load_frames(g):
    n = len([k for k in g.files if k.startswith("frame")])
    return [g["frame%d" % i] for i in range(n)]


def test_native_tracker_reproduces_reference_run(golden_dir):
    g = np.load(os.path.join(golden_dir, "deepsort_tracker.npz"))
    got, _ = run_native(load_frames(g))
    assert got.shape == g["results"].shape
    np.testing.assert_array_equal(got[:, :2], g["results"][:, :2])             # frame, track id: bit-exact
    assert np.abs(got[:, 2:] - g["results"][:, 2:]).max() < 1e-9               # float64 Kalman state


def test_native_tracker_crowded_sequence_ids_and_life_cycle(golden_dir):
    """90 frames: crossings, look-alikes, misses, false positives, tracks ageing out after 60 misses, late arrivals."""
    g = np.load(os.path.join(golden_dir, "deepsort_tracker_crowd.npz"))
    got, life = run_native(load_frames(g))
    assert got.shape == g["results"].shape
    np.testing.assert_array_equal(got[:, :2], g["results"][:, :2])
    assert np.abs(got[:, 2:] - g["results"][:, 2:]).max() < 1e-9
    np.testing.assert_array_equal(life, g["life"])                             # id, state, hits, age, time_since_update
    assert (life[:, 5] > 55).any()                                             # the ageing-out path was exercised


def test_native_tracker_equals_oracle_with_other_parameters(golden_dir):
    """n_init = 3 / max_age = 4 / budget-free gallery exercise tentative deletion and early ageing out (oracle = the pinned
    restatement of the reference loop)."""
    from object_detection_tracking_b200.tracking import Tracker
    g = np.load(os.path.join(golden_dir, "deepsort_tracker_crowd.npz"))
    frames = load_frames(g)[:40]
    m_ref = nn_matching.NearestNeighborDistanceMetric("cosine", 0.3, None)
    ref = deepsort.Tracker(m_ref, max_iou_distance=0.7, max_age=4, n_init=3)
    nat = Tracker(nn_matching.NearestNeighborDistanceMetric("cosine", 0.3, None), max_iou_distance=0.7, max_age=4,
                  n_init=3, cost_fn=oracle_cost)
    for rows in frames:
        dets = [deepsort.Detection(r[:4], r[4], r[5:]) for r in rows]
        ref.predict(); ref.update(dets)
        nat.predict(); nat.update(dets)
        assert [t.track_id for t in nat.tracks] == [t.track_id for t in ref.tracks]
        assert [t.state for t in nat.tracks] == [t.state for t in ref.tracks]
        assert [t.time_since_update for t in nat.tracks] == [t.time_since_update for t in ref.tracks]
        for a, b in zip(nat.tracks, ref.tracks):
            assert np.abs(a.mean - b.mean).max() < 1e-9 and np.abs(a.covariance - b.covariance).max() < 1e-9


def test_empty_frames_and_first_frame():
    from object_detection_tracking_b200.tracking import Tracker
    trk = Tracker(nn_matching.NearestNeighborDistanceMetric("cosine", 0.5, 5), cost_fn=oracle_cost)
    trk.predict()
    trk.update([])                                   # nothing seen yet: no native state is created
    assert trk.tracks == []
    d = deepsort.Detection([10, 20, 30, 60], 0.9, np.ones(8, np.float32))
    trk.predict(); trk.update([d])
    assert len(trk.tracks) == 1 and trk.tracks[0].is_tentative() and trk.tracks[0].track_id == 1
    trk.predict(); trk.update([])                    # a tentative track that is missed is deleted (track.py:147-150)
    assert trk.tracks == []


def test_linear_sum_assignment_equals_scipy_including_ties():
    from scipy.optimize import linear_sum_assignment as sp
    from object_detection_tracking_b200.tracking import linear_sum_assignment
    rng = np.random.default_rng(0)
    for it in range(1200):
        nr, nc = rng.integers(1, 16, 2)
        kind = it % 4
        if kind == 0:
            c = rng.standard_normal((nr, nc))
        elif kind == 1:
            c = rng.integers(0, 4, (nr, nc)).astype(float)                     # many ties
        elif kind == 2:
            c = rng.uniform(0, 1, (nr, nc)); c[c > 0.5] = 0.5 + 1e-5           # thresholded, as min_cost_matching does
        else:
            c = np.full((nr, nc), 0.7)
        r0, c0 = sp(c)
        r1, c1 = linear_sum_assignment(c)
        np.testing.assert_array_equal(r0, r1)
        np.testing.assert_array_equal(c0, c1)
    with pytest.raises(RuntimeError):
        linear_sum_assignment(np.array([[np.nan, 1.0], [1.0, 2.0]]))


def test_pre_tracker_nms_matches_reference_golden(golden_dir):
    from object_detection_tracking_b200.tracking import non_max_suppression
    g = np.load(os.path.join(golden_dir, "track_nms.npz"))
    for case in range(6):
        boxes, scores, thr = g["boxes%d" % case], g["scores%d" % case], float(g["thr%d" % case])
        assert non_max_suppression(boxes, thr, scores) == g["keep_scored%d" % case].tolist()
        assert non_max_suppression(boxes, thr) == g["keep_plain%d" % case].tolist()
    assert non_max_suppression(np.zeros((0, 4)), 0.5) == []


def test_create_obj_infos_matches_reference(golden_dir):
    """Detector output -> tracker input against deep_sort.utils.create_obj_infos run on the same
    arrays (tests/golden/make_golden_tmot.py: obj_infos)."""
    from object_detection_tracking_b200.tracking import create_obj_infos, preprocess_detections
    g = np.load(os.path.join(golden_dir, "obj_infos.npz"))
    id2class = {1: "Person", 2: "Vehicle", 3: "Bike", 4: "car", 5: "person"}
    coco_map = {"car": "Vehicle", "person": "Person"}
    pooled = g["feats"].mean(axis=(2, 3))
    cases = [("actev", ["Person"], 0.6, 0.0, 0.75, False, g["feats"]), ("coco", ["Vehicle"], 0.3, 40.0, 1.5, True, g["feats"]),
             ("pooled", ["Person"], 0.2, 0.0, 1.0, False, pooled)]
    for name, objs, conf, minh, scale, coco, ft in cases:
        boxes = g["boxes"].copy()
        dets = create_obj_infos(7, boxes, g["probs"], g["labels"], ft, id2class, objs, conf, minh, scale,
                                is_coco_model=coco, coco_to_actev_mapping=coco_map)
        assert len(dets) == len(g[name + "_conf"]) > 0
        np.testing.assert_array_equal(np.asarray([d.tlwh for d in dets]).reshape(-1, 4), g[name + "_tlwh"])
        np.testing.assert_array_equal([d.confidence for d in dets], g[name + "_conf"])
        np.testing.assert_allclose(np.asarray([d.feature for d in dets]), g[name + "_feat"], rtol=0, atol=1e-6)
        assert dets[0].tlwh.dtype == np.float64 and dets[0].feature.dtype == np.float32
        np.testing.assert_array_equal(boxes, g["boxes"])              # the caller's boxes are not modified
    tm = preprocess_detections(g["boxes"], g["probs"], g["labels"], g["feats"], id2class, ["Person"], 0.6, 0.75)
    assert len(tm) == len(g["actev_conf"])
    np.testing.assert_array_equal(np.asarray([t[0] for t in tm], dtype=np.float64), g["actev_tlwh"])
    np.testing.assert_array_equal([t[1] for t in tm], g["actev_conf"])


def test_track_table_post_processing_matches_reference(golden_dir):
    """linear_inter_bbox / filter_short_objs (deep_sort/utils.py:47-113) against the reference functions' output."""
    from object_detection_tracking_b200.tracking import filter_short_objs, linear_inter_bbox
    g = np.load(os.path.join(golden_dir, "track_post.npz"))
    inter = linear_inter_bbox(g["data"], 8)
    assert inter.shape == g["inter"].shape and inter.shape[0] > g["data"].shape[0]
    np.testing.assert_allclose(inter, g["inter"], rtol=0, atol=1e-9)
    filt = filter_short_objs(inter)
    assert filt.shape == g["filt"].shape and filt.shape[0] < inter.shape[0]
    np.testing.assert_allclose(filt, g["filt"], rtol=0, atol=1e-9)
    assert linear_inter_bbox(np.zeros((0, 7)), 8).shape == (0, 7) and filter_short_objs(np.zeros((0, 7))).shape == (0, 7)
