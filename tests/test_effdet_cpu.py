"""EfficientDet oracle (oracle/effdet.py) pinned against the numpy halves of the reference's own efficientdet/anchors.py
(tests/golden/effdet_numpy.npz, make_golden.py:effdet_numpy) + structural properties of the post-processing."""
import os

import numpy as np
import pytest
import torch

from object_detection_tracking_b200.effdet_config import BIFPN_NODES, feat_sizes, make_effdet_config
from object_detection_tracking_b200.synth import synth_effdet_weights
from oracle import effdet as oe
from oracle import tf_ops


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "effdet_numpy.npz"))


@pytest.mark.parametrize("tag,h,w,scale", [("a", 256, 384, 4.0), ("b", 512, 640, 5.0)])
def test_anchor_grid_matches_reference(golden, tag, h, w, scale):
    cfg = make_effdet_config("efficientdet-d0", h, w, anchor_scale=scale)
    got = oe.anchor_boxes(cfg, feat_sizes(cfg))
    ref = golden["anchors_" + tag]
    assert got.dtype == np.float32 and got.shape == ref.shape
    np.testing.assert_array_equal(got, ref.astype(np.float32))


def test_decode_and_sigmoid_match_reference(golden):
    got = oe.decode_boxes(golden["dec_codes"], golden["dec_anchors"])
    np.testing.assert_array_equal(got, golden["dec_boxes"].astype(np.float32))
    np.testing.assert_array_equal(oe.sigmoid(golden["sig_logits"]), golden["sig_scores"].astype(np.float32))


def test_maxpool_same_padding():
    x = torch.arange(2 * 5 * 6, dtype=torch.float32).reshape(1, 2, 5, 6)
    y = oe.maxpool_same(x, 2)                      # odd height: TF SAME pads (1, 1) rows; even width: (0, 1)
    assert tuple(y.shape) == (1, 2, 3, 3)
    xn = x.numpy()[0]
    for c in range(2):
        for i in range(3):
            for j in range(3):
                r0, c0 = 2 * i - 1, 2 * j
                win = xn[c, max(r0, 0):min(r0 + 3, 5), c0:min(c0 + 3, 6)]
                assert y[0, c, i, j].item() == win.max()


def test_bifpn_node_table_is_the_reference_topology():
    # efficientdet_arch.py:508-522: 8 nodes per cell, top-down then bottom-up, last node of each level is the output
    assert [n[0] for n in BIFPN_NODES] == [6, 5, 4, 3, 4, 5, 6, 7]
    assert BIFPN_NODES[4][1] == (1, 7, 8) and BIFPN_NODES[7][1] == (4, 11)


@pytest.fixture(scope="module")
def tiny():
    cfg = make_effdet_config("efficientdet-d0", 128, 256, fpn_cell_repeats=2, box_class_repeats=1)
    W = synth_effdet_weights(cfg)
    rng = np.random.default_rng(5)
    fs = feat_sizes(cfg)
    feats = {l: np.abs(rng.standard_normal((cfg.backbone_channels[l - 3],) + fs[l])).astype(np.float32) for l in (3, 4, 5)}
    return cfg, W, feats, oe.forward_from_features(cfg, W, feats, image_scale=1.5, stages=True)


def test_forward_shapes_and_postprocess_properties(tiny):
    cfg, W, feats, r = tiny
    fs = feat_sizes(cfg)
    for l in range(3, 8):
        assert r["fpn"][l].shape == (cfg.fpn_num_filters,) + fs[l]
        assert r["cls_out"][l].shape == fs[l] + (9 * 90,) and r["box_out"][l].shape == fs[l] + (36,)
    n = len(r["final_probs"])
    assert 0 < n <= cfg.result_per_im
    assert np.all(np.diff(r["final_probs"]) <= 0) and np.all(r["final_probs"] > cfg.result_score_thres)
    assert r["final_labels"].min() >= 1 and r["final_labels"].max() <= 90
    assert set(np.unique(r["levels"])) <= set(range(3, 8))
    assert r["fpn_box_feat"].shape == (n, cfg.fpn_num_filters)
    b = r["final_boxes"] / np.float32(1.5)                      # back to network coords, class-agnostic NMS @0.5
    yx = b[:, [1, 0, 3, 2]]
    keep = tf_ops.non_max_suppression(yx, r["final_probs"], n, cfg.nms_iou_threshold)
    assert len(keep) == n                                        # idempotent: nothing left to suppress


def test_topk_is_global_over_levels_and_classes(tiny):
    cfg, W, feats, r = tiny
    flat = np.concatenate([r["cls_out"][l].reshape(-1) for l in range(3, 8)])
    kth = np.sort(flat)[::-1][min(cfg.max_detection_topk, flat.size) - 1]
    logit = np.log(r["final_probs"].astype(np.float64) / (1 - r["final_probs"].astype(np.float64)))
    assert np.all(logit >= kth - 1e-4)


def test_fast_attention_weights_are_used(tiny):
    cfg, W, feats, r = tiny
    W2 = dict(W)
    W2["fpn_cells/cell_0/fnode0/WSM"] = np.float32(-1.0)         # relu -> 0: the first edge drops out of the node
    r2 = oe.forward_from_features(cfg, W2, feats, image_scale=1.5, stages=True)
    assert np.abs(r2["fpn"][6] - r["fpn"][6]).max() > 1e-3


# ---- EfficientNet backbone geometry + pre-processing oracle -----------------------------------------------------------
def test_filter_and_repeat_rounding_match_reference(golden):
    from oracle import efficientnet as on
    names = sorted(on.PARAMS)
    base_f, base_r = [32, 16, 24, 40, 80, 112, 192, 320, 1280], [1, 2, 3, 4]
    for row, name in enumerate(names):
        width, depth = on.PARAMS[name]
        assert [on.round_filters(f, width) for f in base_f] == list(golden["round_filters"][row])
        assert [on.round_repeats(r, depth) for r in base_r] == list(golden["round_repeats"][row])


def test_backbone_endpoints_match_the_detector_table():
    from object_detection_tracking_b200.effdet_config import _TABLE, BACKBONE_OF, efficientnet_blocks
    from oracle import efficientnet as on
    for det, bb in BACKBONE_OF.items():
        stem, blocks = on.block_specs(bb)
        red = on.endpoint_blocks(blocks)
        assert len(red) == 5
        assert tuple(blocks[red[k]].cout for k in (2, 3, 4)) == _TABLE[det][5]
        s2, b2 = efficientnet_blocks(bb)                              # the product package's own copy of the geometry
        assert s2 == stem and [(b.kernel, b.stride, b.expand, b.cin, b.cout) for b in b2] == \
            [(b.kernel, b.stride, b.expand, b.cin, b.cout) for b in blocks]
    assert len(on.block_specs("efficientnet-b6")[1]) == 45


def test_preprocess_resize_pad_and_scale():
    from oracle import efficientnet as on
    rng = np.random.default_rng(0)
    frame = rng.integers(0, 256, (300, 500, 3)).astype(np.uint8)
    img, scale = on.preprocess(frame, 256, 384)
    assert img.shape == (256, 384, 3) and img.dtype == np.float32
    assert abs(scale - 500 / 384) < 1e-6                              # width-limited: image_scale_to_original
    sh = int(np.float32(300) * (np.float32(384) / np.float32(500)))
    assert not img[sh:].any() and img[:sh].any()                      # zero padding below the resized frame
    # pixel (0,0) maps to source (0,0): RGB order, /255, mean/std
    exp = (frame[0, 0, ::-1].astype(np.float32) * np.float32(1 / 255.0) - np.array([0.485, 0.456, 0.406], np.float32)) \
        / np.array([0.229, 0.224, 0.225], np.float32)
    np.testing.assert_allclose(img[0, 0], exp, rtol=1e-6)
    same, s1 = on.preprocess(frame[:256, :384], 256, 384)             # identity size: no interpolation
    assert s1 == 1.0
    np.testing.assert_allclose(same[5, 7], (frame[5, 7, ::-1].astype(np.float32) * np.float32(1 / 255.0)
                               - np.array([0.485, 0.456, 0.406], np.float32)) / np.array([0.229, 0.224, 0.225], np.float32), rtol=1e-6)


def test_same_padding_and_backbone_shapes():
    from object_detection_tracking_b200.synth import synth_efficientnet_weights
    from oracle import efficientnet as on
    x = torch.zeros(1, 1, 7, 10)
    assert tuple(on.same_pad(x, 3, 2).shape) == (1, 1, 9, 11)         # 7 -> out 4: total 2 (1,1); 10 -> out 5: total 1 (0,1)
    assert tuple(on.same_pad(x, 5, 1).shape) == (1, 1, 11, 14)
    W = synth_efficientnet_weights("efficientnet-b0")
    img = np.random.default_rng(1).standard_normal((128, 256, 3)).astype(np.float32)
    r = on.forward(img, W, "efficientnet-b0")
    assert r[3].shape == (40, 16, 32) and r[4].shape == (112, 8, 16) and r[5].shape == (320, 4, 8)
    assert all(np.isfinite(r[l]).all() for l in (3, 4, 5))


def test_coco_tables(golden):
    from object_detection_tracking_b200 import class_ids as c
    assert sorted(c.coco_id_mapping) == list(golden["coco_ids"])
    assert [c.coco_id_mapping[i] for i in golden["coco_ids"]] == [str(n) for n in golden["coco_names"]]
    assert c.effdet_labels_to_coco80(golden["coco_ids"]) == list(golden["coco_dense"])
    assert c.coco_id_mapping[1] == "person" and c.coco_id_mapping[90] == "toothbrush" and 12 not in c.coco_id_mapping
    assert c.coco_id_mapping[13] == "stop sign" and c.coco_id_mapping[67] == "dining table"
    assert c.effdet_labels_to_coco80([1, 3, 90]) == [1, 3, 80]
    assert c.coco_id_mapping_reverse["car"] == 3
