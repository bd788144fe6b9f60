"""Pins the anchor restatement (oracle/anchors.py) to the reference: golden vectors produced by the
reference's own generate_anchors.py (tests/golden/make_golden.py) and its known-answer comment
(generate_anchors.py:30-38, 0-indexed = minus 1)."""
import os

import numpy as np

from oracle.anchors import generate_anchors, get_all_anchors


def test_default_anchors_known_answer():
    expect = np.array([[-84, -40, 99, 55], [-176, -88, 191, 103], [-360, -184, 375, 199],
                       [-56, -56, 71, 71], [-120, -120, 135, 135], [-248, -248, 263, 263],
                       [-36, -80, 51, 95], [-80, -168, 95, 183], [-168, -344, 183, 359]], dtype=np.float64)
    np.testing.assert_array_equal(generate_anchors(), expect)


def test_fpn_cell_anchors_match_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "anchors.npz"))
    np.testing.assert_array_equal(generate_anchors(), g["default"])
    for stride, size in zip((4, 8, 16, 32, 64), (32, 64, 128, 256, 512)):
        a = generate_anchors(stride, ratios=(0.5, 1, 2), scales=np.array([size], dtype=np.float64) / stride)
        np.testing.assert_array_equal(a, g["s%d" % stride])
    # SURVEY.md 8c vectors
    np.testing.assert_array_equal(g["s4"], [[-22, -10, 25, 13], [-14, -14, 17, 17], [-10, -22, 13, 25]])
    np.testing.assert_array_equal(g["s64"], [[-332, -152, 395, 215], [-224, -224, 287, 287], [-148, -328, 211, 391]])


def test_get_all_anchors_field():
    f = get_all_anchors(16, [128], (0.5, 1, 2), 1280)
    assert f.shape == (80, 80, 3, 4) and f.dtype == np.float32
    # shift k*stride and +1 on x2,y2 (utils.py:633-657)
    np.testing.assert_array_equal(f[0, 0], [[-84, -40, 100, 56], [-56, -56, 72, 72], [-36, -80, 52, 96]])
    np.testing.assert_array_equal(f[2, 5] - f[0, 0], np.tile([80, 32, 80, 32], (3, 1)))


def test_anchor_count_720x1280():
    from object_detection_tracking_b200.config import backbone_geometry, make_config
    cfg = make_config()
    geo = backbone_geometry(720, 1280, cfg)
    assert geo["c1"] == (368, 640) and geo["c"] == [(184, 320), (92, 160), (46, 80), (23, 40)]
    assert geo["p"] == [(180, 320), (90, 160), (45, 80), (23, 40), (12, 20)]
    assert sum(h * w * 3 for h, w in geo["p"]) == 230280
