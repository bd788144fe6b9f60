"""ORACLE (test infrastructure only).  The reference's DeepSORT loop restated in one module so that the
GPU appearance metric can be checked for what it is used for: track identities.

Follows deep_sort/{detection,kalman_filter,track,iou_matching,linear_assignment,tracker}.py of the reference
(this fork's constants: Tracker(max_iou_distance=0.5, max_age=60, n_init=1), tracker.py:40).  Pinned
against a run of the reference's own Tracker (tests/golden/deepsort_tracker.npz, produced by
tests/golden/make_golden.py) -- identical ids and boxes, see tests/test_oracle_ops.py.
All state is float64 as in the reference; detection features are float32 (detection.py:30-32).
"""
from __future__ import annotations

import numpy as np
import scipy.linalg
from scipy.optimize import linear_sum_assignment

CHI2INV95_4 = 9.4877        # kalman_filter.py:11-20, 4 degrees of freedom
INFTY_COST = 1e+5           # linear_assignment.py:9
TENTATIVE, CONFIRMED, DELETED = 1, 2, 3


class Detection(object):    # detection.py:27-49
    def __init__(self, tlwh, confidence, feature):
        self.tlwh = np.asarray(tlwh, dtype=np.float64)
        self.confidence = float(confidence)
        self.feature = np.asarray(feature, dtype=np.float32)

    def to_xyah(self):
        r = self.tlwh.copy()
        r[:2] += r[2:] / 2
        r[2] /= r[3]
        return r


class Kalman(object):
    """kalman_filter.py:23-232: constant-velocity model on (x, y, a, h)."""
    F = np.eye(8)
    F[:4, 4:] = np.eye(4)       # dt = 1
    Hm = np.eye(4, 8)
    wp, wv = 1. / 20, 1. / 160

    def initiate(self, z):                                        # :55-87
        mean = np.r_[z, np.zeros_like(z)]
        h = z[3]
        std = [2 * self.wp * h, 2 * self.wp * h, 1e-2, 2 * self.wp * h,
               10 * self.wv * h, 10 * self.wv * h, 1e-5, 10 * self.wv * h]
        return mean, np.diag(np.square(std))

    def predict(self, mean, cov):                                 # :89-124
        h = mean[3]
        q = np.diag(np.square(np.r_[[self.wp * h, self.wp * h, 1e-2, self.wp * h],
                                    [self.wv * h, self.wv * h, 1e-5, self.wv * h]]))
        return np.dot(self.F, mean), np.linalg.multi_dot((self.F, cov, self.F.T)) + q

    def project(self, mean, cov):                                 # :126-154
        h = mean[3]
        r = np.diag(np.square([self.wp * h, self.wp * h, 1e-1, self.wp * h]))
        return np.dot(self.Hm, mean), np.linalg.multi_dot((self.Hm, cov, self.Hm.T)) + r

    def update(self, mean, cov, z):                               # :156-190
        pm, pc = self.project(mean, cov)
        chol, lower = scipy.linalg.cho_factor(pc, lower=True, check_finite=False)
        gain = scipy.linalg.cho_solve((chol, lower), np.dot(cov, self.Hm.T).T, check_finite=False).T
        return mean + np.dot(z - pm, gain.T), cov - np.linalg.multi_dot((gain, pc, gain.T))

    def gating_distance(self, mean, cov, zs):                     # :192-232
        pm, pc = self.project(mean, cov)
        L = np.linalg.cholesky(pc)
        y = scipy.linalg.solve_triangular(L, (zs - pm).T, lower=True, check_finite=False, overwrite_b=True)
        return np.sum(y * y, axis=0)


class Track(object):        # track.py:19-166
    def __init__(self, mean, cov, track_id, n_init, max_age, feature):
        self.mean, self.covariance, self.track_id = mean, cov, track_id
        self.hits, self.age, self.time_since_update = 1, 1, 0
        self.state = TENTATIVE
        self.features = [feature] if feature is not None else []
        self.n_init, self.max_age = n_init, max_age

    def to_tlwh(self):
        r = self.mean[:4].copy()
        r[2] *= r[3]
        r[:2] -= r[2:] / 2
        return r

    def is_confirmed(self):
        return self.state == CONFIRMED


def _iou_cost(tracks, dets, ti, di):
    """iou_matching.py:42-81."""
    cost = np.zeros((len(ti), len(di)))
    cand = np.asarray([dets[i].tlwh for i in di])
    for r, k in enumerate(ti):
        if tracks[k].time_since_update > 1:
            cost[r, :] = INFTY_COST
            continue
        b = tracks[k].to_tlwh()
        tl = np.maximum(b[:2], cand[:, :2])
        br = np.minimum(b[:2] + b[2:], cand[:, :2] + cand[:, 2:])
        inter = np.maximum(0., br - tl).prod(axis=1)
        cost[r, :] = 1. - inter / (b[2:].prod() + cand[:, 2:].prod(axis=1) - inter)
    return cost


def _min_cost_matching(metric_fn, max_distance, tracks, dets, ti, di):
    """linear_assignment.py:12-78."""
    if len(di) == 0 or len(ti) == 0:
        return [], ti, di
    cost = metric_fn(tracks, dets, ti, di)
    cost[cost > max_distance] = max_distance + 1e-5
    rows, cols = linear_sum_assignment(cost)
    matches, ut, ud = [], [], []
    for c, d in enumerate(di):
        if c not in cols:
            ud.append(d)
    for r, t in enumerate(ti):
        if r not in rows:
            ut.append(t)
    for r, c in zip(rows, cols):
        if cost[r, c] > max_distance:
            ut.append(ti[r]); ud.append(di[c])
        else:
            matches.append((ti[r], di[c]))
    return matches, ut, ud


class Tracker(object):
    """tracker.py:10-138."""

    def __init__(self, metric, max_iou_distance=0.5, max_age=60, n_init=1):
        self.metric, self.max_iou_distance, self.max_age, self.n_init = metric, max_iou_distance, max_age, n_init
        self.kf = Kalman()
        self.tracks = []
        self._next_id = 1

    def predict(self):                                            # :50-56, track.py:107-124
        for t in self.tracks:
            t.mean, t.covariance = self.kf.predict(t.mean, t.covariance)
            t.age += 1
            t.time_since_update += 1

    def _gated(self, tracks, dets, ti, di):                       # :94-104, linear_assignment.py:148-194
        feats = np.array([dets[i].feature for i in di])
        cost = self.metric.distance(feats, np.array([tracks[i].track_id for i in ti]))
        zs = np.asarray([dets[i].to_xyah() for i in di])
        for r, k in enumerate(ti):
            g = self.kf.gating_distance(tracks[k].mean, tracks[k].covariance, zs)
            cost[r, g > CHI2INV95_4] = INFTY_COST
        return cost

    def _match(self, dets):                                       # :92-131
        confirmed = [i for i, t in enumerate(self.tracks) if t.is_confirmed()]
        unconfirmed = [i for i, t in enumerate(self.tracks) if not t.is_confirmed()]
        # matching cascade (linear_assignment.py:81-145)
        ud = list(range(len(dets)))
        matches_a = []
        for level in range(self.max_age):
            if len(ud) == 0:
                break
            lvl = [k for k in confirmed if self.tracks[k].time_since_update == 1 + level]
            if not lvl:
                continue
            m, _, ud = _min_cost_matching(self._gated, self.metric.matching_threshold, self.tracks, dets, lvl, ud)
            matches_a += m
        ut_a = list(set(confirmed) - set(k for k, _ in matches_a))
        iou_cand = unconfirmed + [k for k in ut_a if self.tracks[k].time_since_update == 1]
        ut_a = [k for k in ut_a if self.tracks[k].time_since_update != 1]
        matches_b, ut_b, ud = _min_cost_matching(_iou_cost, self.max_iou_distance, self.tracks, dets, iou_cand, ud)
        return matches_a + matches_b, list(set(ut_a + ut_b)), ud

    def update(self, dets):                                       # :57-90
        matches, ut, ud = self._match(dets)
        for k, d in matches:                                      # track.py:126-145
            t = self.tracks[k]
            t.mean, t.covariance = self.kf.update(t.mean, t.covariance, dets[d].to_xyah())
            t.features.append(dets[d].feature)
            t.hits += 1
            t.time_since_update = 0
            if t.state == TENTATIVE and t.hits >= t.n_init:
                t.state = CONFIRMED
        for k in ut:                                              # track.py:147-153
            t = self.tracks[k]
            if t.state == TENTATIVE or t.time_since_update > t.max_age:
                t.state = DELETED
        for d in ud:                                              # :133-138
            mean, cov = self.kf.initiate(dets[d].to_xyah())
            self.tracks.append(Track(mean, cov, self._next_id, self.n_init, self.max_age, dets[d].feature))
            self._next_id += 1
        self.tracks = [t for t in self.tracks if t.state != DELETED]
        active = [t.track_id for t in self.tracks if t.is_confirmed()]
        feats, targets = [], []
        for t in self.tracks:
            if not t.is_confirmed():
                continue
            feats += t.features
            targets += [t.track_id for _ in t.features]
            t.features = []
        self.metric.partial_fit(np.asarray(feats), np.asarray(targets), active)


def run_sequence(frames, metric):
    """frames: list of [n, 5 + D] arrays (tlwh, confidence, feature) -> rows (frame, id, tlwh) of the
    confirmed, recently updated tracks, as tests/golden/make_golden.py records them."""
    tracker = Tracker(metric)
    out = []
    for f, rows in enumerate(frames):
        dets = [Detection(r[:4], r[4], r[5:]) for r in rows]
        tracker.predict()
        tracker.update(dets)
        for t in tracker.tracks:
            if t.is_confirmed() and t.time_since_update <= 1:
                out.append([f, t.track_id] + t.to_tlwh().tolist())
    return np.asarray(out, dtype=np.float64).reshape(-1, 6)
