"""Checkpoint importer, host half (reference: initialize(), obj_detect_tracking.py:392-448): the
variable manifest in the reference's checkpoint naming, ':0' tensor names, ignored extras, loud failure on missing /
mis-shaped variables (the reference would silently keep random initial values)."""
import numpy as np
import pytest

from object_detection_tracking_b200.backend import check_weights
from object_detection_tracking_b200.config import make_config
from object_detection_tracking_b200.synth import frcnn_weight_shapes, synth_weights


def test_manifest_equals_synthetic_checkpoint_and_reference_param_counts():
    cfg = make_config(resnet_num_block=(1, 1, 2, 1))
    W = synth_weights(cfg, 3)
    sh = frcnn_weight_shapes(cfg)
    assert set(sh) == set(W)
    for k in sh:
        assert tuple(W[k].shape) == tuple(sh[k]), k
    # full-size graph: ResNet-101 convs 42.39 M + FPN 3.34 M + RPN 0.59 M = 46.33 M kernel weights (SURVEY 8's 48.7 M also
    # figure is an estimate that also counts BatchNorm vectors and biases), 14.0 M fc for 15 classes
    r101 = frcnn_weight_shapes(make_config())
    conv = sum(int(np.prod(s)) for n, s in r101.items() if n.endswith("/W") and len(s) == 4)
    fc = sum(int(np.prod(s)) for n, s in r101.items() if n.endswith("/W") and len(s) == 2)
    assert abs(conv - 46.33e6) < 0.02e6 and abs(fc - 14.0e6) < 0.1e6
    assert len([n for n in r101 if n.endswith("/conv2/W")]) == 33 and "group2/block22/conv3/bn/variance/EMA" in r101
    r50 = frcnn_weight_shapes(make_config(resnet50=True, num_class=81))
    assert "group2/block5/conv1/W" in r50 and "group2/block6/conv1/W" not in r50
    assert r50["fastrcnn/outputs/box/W"] == (1024, 324)
    agn = frcnn_weight_shapes(make_config(use_frcnn_class_agnostic=True))
    assert agn["fastrcnn/outputs/box/W"] == (1024, 4)


def test_importer_accepts_tensor_names_and_ignores_extras():
    cfg = make_config(resnet_num_block=(1, 1, 1, 1))
    W = synth_weights(cfg, 4)
    ck = {k + ":0": v.astype(np.float64) for k, v in W.items()}          # tensor names, other dtype
    ck["global_step:0"] = np.int64(90000)
    ck["learning_rate"] = np.float32(0.001)
    ck["conv0/W/Momentum:0"] = np.zeros_like(W["conv0/W"])
    ck["maskrcnn/fcn0/W:0"] = np.zeros((3, 3, 256, 256), np.float32)      # head this graph does not build
    got = check_weights(cfg, ck)
    assert set(got) == set(W)
    for k in W:
        assert got[k].dtype == np.float32 and got[k].flags["C_CONTIGUOUS"]
        np.testing.assert_array_equal(got[k], W[k])


def test_importer_fails_loudly_on_missing_or_misshaped():
    cfg = make_config(resnet_num_block=(1, 1, 1, 1))
    W = synth_weights(cfg, 5)
    bad = dict(W)
    del bad["group3/block0/conv2/bn/mean/EMA"]
    bad["fastrcnn/outputs/class/W"] = np.zeros((1024, 81), np.float32)    # COCO head into a 15-class graph
    with pytest.raises(ValueError) as e:
        check_weights(cfg, bad)
    assert "group3/block0/conv2/bn/mean/EMA" in str(e.value) and "fastrcnn/outputs/class/W" in str(e.value)
    with pytest.raises(ValueError):
        check_weights(make_config(), W)                                    # R101 graph, 4-block checkpoint


def test_model_load_npz_roundtrip(tmp_path):
    from object_detection_tracking_b200.backend import get_model
    cfg = make_config(resnet_num_block=(1, 1, 1, 1))
    W = synth_weights(cfg, 6)
    path = str(tmp_path / "ckpt.npz")
    np.savez(path, **{k + ":0": v for k, v in W.items()})
    model = get_model(cfg, gpuid=0)
    model.load_npz(path)                                                   # host side only: no device needed yet
    assert set(model._weights) == set(W)
    np.testing.assert_array_equal(model._weights["rpn/box/b"], W["rpn/box/b"])


def test_mask_head_variables_join_the_manifest_with_add_mask():
    cfg = make_config(resnet_num_block=(1, 1, 1, 1), add_mask=True)
    sh = frcnn_weight_shapes(cfg)
    assert sh["maskrcnn/fcn3/W"] == (3, 3, 256, 256) and sh["maskrcnn/deconv/W"] == (2, 2, 256, 256)
    assert sh["maskrcnn/conv/W"] == (1, 1, 256, 14) and sh["maskrcnn/conv/b"] == (14,)
    W = synth_weights(cfg, 9)
    assert set(W) == set(sh) and all(tuple(W[k].shape) == tuple(sh[k]) for k in sh)
    base = synth_weights(make_config(resnet_num_block=(1, 1, 1, 1)), 9)
    assert "maskrcnn/conv/W" not in base
    np.testing.assert_array_equal(base["fastrcnn/fc7/W"], W["fastrcnn/fc7/W"])      # detector weights independent of add_mask
    with pytest.raises(ValueError):
        check_weights(cfg, base)                                                      # mask graph, checkpoint without the head
