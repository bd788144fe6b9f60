"""CPU checks of the TF-op restatements (oracle/tf_ops.py, oracle/frcnn.py) against independent
implementations (torchvision) and hand-computed cases; and of the DeepSORT metric restatement against
golden vectors produced by the reference's own code."""
import os

import numpy as np
import pytest
import torch

from oracle import frcnn, nn_matching, tf_ops


def test_nms_known_answer():
    boxes = np.array([[0, 0, 10, 10], [1, 1, 11, 11], [20, 20, 30, 30], [0, 0, 10, 10.5]], dtype=np.float32)
    scores = np.array([0.9, 0.8, 0.7, 0.9], dtype=np.float32)
    # ties -> lower index first; IoU(0,1) = 81/119 = 0.68 > 0.5 suppressed; box 3 IoU with 0 = 100/105
    keep = tf_ops.non_max_suppression(boxes, scores, 10, 0.5)
    assert keep.tolist() == [0, 2]
    keep = tf_ops.non_max_suppression(boxes, scores, 10, 0.7)
    assert keep.tolist() == [0, 1, 2]
    assert tf_ops.non_max_suppression(boxes, scores, 1, 0.7).tolist() == [0]
    # zero-area boxes never suppress / are never suppressed (IoU 0)
    z = np.array([[5, 5, 5, 9], [0, 0, 10, 10]], dtype=np.float32)
    assert tf_ops.non_max_suppression(z, np.array([1.0, 0.5], np.float32), 10, 0.1).tolist() == [0, 1]


def test_nms_matches_torchvision_on_random_boxes():
    tv = pytest.importorskip("torchvision")
    rng = np.random.default_rng(0)
    xy = rng.uniform(0, 200, (300, 2)); wh = rng.uniform(5, 80, (300, 2))
    boxes = np.concatenate([xy, xy + wh], 1).astype(np.float32)
    scores = rng.uniform(0, 1, 300).astype(np.float32)
    ours = tf_ops.non_max_suppression(boxes, scores, 300, 0.5)
    ref = tv.ops.nms(torch.from_numpy(boxes), torch.from_numpy(scores), 0.5).numpy()
    assert ours.tolist() == ref.tolist()


def test_top_k_order():
    s = np.array([0.1, 0.9, 0.5, 0.9, -1.0], dtype=np.float32)
    v, i = tf_ops.top_k(s, 3)
    assert i.tolist() == [1, 3, 2] and v.tolist() == [np.float32(0.9), np.float32(0.9), np.float32(0.5)]


def test_roi_align_matches_torchvision_aligned():
    tv = pytest.importorskip("torchvision")
    rng = np.random.default_rng(1)
    feat = rng.standard_normal((8, 40, 56)).astype(np.float32)
    xy = rng.uniform(4, 15, (20, 2)); wh = rng.uniform(3, 20, (20, 2))
    boxes = np.concatenate([xy, xy + wh], 1).astype(np.float32)     # interior boxes (no extrapolation)
    ours = frcnn.roi_align(feat, boxes, 7)
    rois = torch.cat([torch.zeros(20, 1), torch.from_numpy(boxes)], 1)
    ref = tv.ops.roi_align(torch.from_numpy(feat)[None], rois, 7, spatial_scale=1.0, sampling_ratio=2, aligned=True)
    np.testing.assert_allclose(ours, ref.numpy(), atol=2e-5, rtol=1e-5)


def test_crop_and_resize_extrapolation_is_zero():
    img = np.ones((4, 4, 1), dtype=np.float32)
    out = tf_ops.crop_and_resize(img, np.array([[-0.5, -0.5, 0.5, 0.5]], np.float32), 4)
    assert out[0, 0, 0, 0] == 0 and out[0, -1, -1, 0] == 1


def test_decode_and_clip():
    anchors = np.array([[0, 0, 16, 16]], dtype=np.float32)
    d = frcnn.decode_bbox_target(np.array([[0.5, -0.25, np.log(2), 100.0]], np.float32), anchors, np.log(1280 / 16.0))
    np.testing.assert_allclose(d, [[0.0, -636.0, 32.0, 644.0]], rtol=1e-5, atol=1e-3)
    c = frcnn.clip_boxes(d, (600, 1280))
    np.testing.assert_allclose(c, [[0.0, 0.0, 32.0, 600.0]], rtol=1e-5, atol=1e-3)


def test_fpn_level_mapping():
    b = np.array([[0, 0, 10, 10], [0, 0, 224, 224], [0, 0, 112, 112], [0, 0, 900, 900], [0, 0, 448, 448]], np.float32)
    assert frcnn.fpn_map_rois_to_levels(b).tolist() == [2, 4, 3, 5, 5]


def test_nn_matching_restatement_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "deepsort_cosine.npz"))
    seg = g["seg"]
    m = nn_matching.NearestNeighborDistanceMetric("cosine", 0.5, 5)
    m.samples = {t: list(g["gallery"][seg[t]:seg[t + 1]]) for t in range(len(seg) - 1)}
    cost = m.distance(g["dets"], list(range(len(seg) - 1)))
    np.testing.assert_array_equal(cost, g["cost"])      # same numpy ops -> bit-exact
    assert cost.dtype == np.float64


def test_nn_matching_euclidean_restatement_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "deepsort_euclid.npz"))
    m = nn_matching.NearestNeighborDistanceMetric("euclidean", 0.3, int(g["budget"]))
    m.partial_fit(g["feats"], g["targets"], list(range(int(g["T"]))))
    cost = m.distance(g["dets"], g["order"].tolist())
    np.testing.assert_array_equal(cost, g["cost"])


def test_oracle_forward_small_is_deterministic_and_shaped():
    from object_detection_tracking_b200.config import make_config
    from object_detection_tracking_b200.synth import synth_frame, synth_weights
    cfg = make_config(resnet_num_block=(1, 1, 1, 1), max_size=160, short_edge_size=96)
    W = synth_weights(cfg, 3)
    img = synth_frame(96, 160, 1).astype(np.float32)
    a = frcnn.forward(cfg, W, img)
    b = frcnn.forward(cfg, W, img)
    assert a["c2345"][0].shape == (256, 24, 40) and a["c2345"][3].shape == (2048, 3, 5)
    assert a["p23456"][4].shape == (256, 2, 3)
    R = a["final_boxes"].shape[0]
    assert 0 < R <= 100 and a["fpn_box_feat"].shape == (R, 256, 7, 7) and a["final_labels"].dtype == np.int64
    np.testing.assert_array_equal(a["final_boxes"], b["final_boxes"])
    assert np.all(np.diff(a["final_probs"]) <= 0)
    assert a["final_boxes"][:, [0, 2]].max() <= 160 and a["final_boxes"][:, [1, 3]].max() <= 96


def _golden_frames(golden_dir):
    g = np.load(os.path.join(golden_dir, "deepsort_tracker.npz"))
    n = len([k for k in g.files if k.startswith("frame")])
    return [g["frame%d" % i] for i in range(n)], g["results"]


def test_deepsort_restatement_reproduces_reference_tracker_run(golden_dir):
    """oracle/deepsort.py vs the reference's Tracker (golden run): identical ids, identical boxes."""
    from oracle import deepsort
    frames, expect = _golden_frames(golden_dir)
    got = deepsort.run_sequence(frames, nn_matching.NearestNeighborDistanceMetric("cosine", 0.5, 5))
    assert got.shape == expect.shape
    np.testing.assert_array_equal(got[:, :2], expect[:, :2])           # frame, track id: bit-exact
    np.testing.assert_allclose(got[:, 2:], expect[:, 2:], rtol=0, atol=1e-9)
    assert len(set(expect[:, 1].tolist())) >= 5                         # the fixture really tracks several objects


def test_mask_head_oracle_deconv_against_explicit_sum():
    """maskrcnn_up4conv_head (models.py:1173-1199) in the oracle: the Conv2DTranspose 2x2 / stride 2 step (TF kernel layout
    [kh, kw, out, in], nn.py:402-412) against the definition out[2y+dy, 2x+dx, o] = sum_i x[y, x, i] W[dy, dx, o, i] + b[o],
    and the whole head's shapes / class selection."""
    import numpy as np
    from types import SimpleNamespace
    from oracle import frcnn
    rng = np.random.default_rng(12)
    md, ncls, R = 16, 5, 3
    W = {}
    for k in range(4):
        W["maskrcnn/fcn%d/W" % k] = (rng.standard_normal((3, 3, md, md)) * 0.1).astype(np.float32)
        W["maskrcnn/fcn%d/b" % k] = (rng.standard_normal(md) * 0.1).astype(np.float32)
    W["maskrcnn/deconv/W"] = (rng.standard_normal((2, 2, md, md)) * 0.2).astype(np.float32)
    W["maskrcnn/deconv/b"] = (rng.standard_normal(md) * 0.1).astype(np.float32)
    W["maskrcnn/conv/W"] = (rng.standard_normal((1, 1, md, ncls - 1)) * 0.3).astype(np.float32)
    W["maskrcnn/conv/b"] = (rng.standard_normal(ncls - 1) * 0.1).astype(np.float32)
    feat = rng.standard_normal((R, md, 14, 14)).astype(np.float32)
    logits = frcnn.maskrcnn_head(feat, W, ncls)
    assert logits.shape == (R, ncls - 1, 28, 28)
    # independent evaluation in float64 numpy
    x = feat.astype(np.float64)
    for k in range(4):
        xp = np.pad(x, ((0, 0), (0, 0), (1, 1), (1, 1)))
        w = W["maskrcnn/fcn%d/W" % k].astype(np.float64)
        y = np.zeros((R, md, 14, 14))
        for dy in range(3):
            for dx in range(3):
                y += np.einsum("rihw,io->rohw", xp[:, :, dy:dy + 14, dx:dx + 14], w[dy, dx])
        x = np.maximum(y + W["maskrcnn/fcn%d/b" % k].astype(np.float64)[None, :, None, None], 0)
    up = np.zeros((R, md, 28, 28))
    wd = W["maskrcnn/deconv/W"].astype(np.float64)
    for dy in range(2):
        for dx in range(2):
            up[:, :, dy::2, dx::2] = np.einsum("rihw,oi->rohw", x, wd[dy, dx])
    up = np.maximum(up + W["maskrcnn/deconv/b"].astype(np.float64)[None, :, None, None], 0)
    ref = np.einsum("rihw,io->rohw", up, W["maskrcnn/conv/W"][0, 0].astype(np.float64)) + W["maskrcnn/conv/b"][None, :, None, None]
    assert np.abs(logits - ref).max() < 1e-4
