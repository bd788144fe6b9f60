"""The reference-facing call surface (get_model / get_feed_dict_forward / sess.run) on the GPU."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_single_image_surface_like_obj_detect_tracking():
    from object_detection_tracking_b200 import backend
    from object_detection_tracking_b200.config import make_config
    from object_detection_tracking_b200.synth import synth_frame, synth_weights
    from oracle import frcnn
    cfg = make_config(resnet_num_block=(1, 1, 1, 1), max_size=256, short_edge_size=192)
    model = backend.get_model(cfg, gpuid=0, controller="/cpu:0")
    Wt = synth_weights(cfg, 7)
    model.set_weights(Wt)
    sess = backend.Session()
    im = synth_frame(192, 256, 2).astype("float32")                       # obj_detect_tracking.py:597
    feed_dict = model.get_feed_dict_forward(im)                           # :610
    sess_input = [model.final_boxes, model.final_labels, model.final_probs, model.fpn_box_feat]
    final_boxes, final_labels, final_probs, box_feats = sess.run(sess_input, feed_dict=feed_dict)   # :632-635
    assert len(box_feats) == len(final_boxes)                             # :648
    assert final_labels.dtype == np.int64 and final_boxes.dtype == np.float32
    assert box_feats.shape[1:] == (256, 7, 7)
    ref = frcnn.forward(cfg, Wt, im, stages=False)
    np.testing.assert_array_equal(final_labels, ref["final_labels"])
    assert np.abs(final_boxes - ref["final_boxes"]).max() < 1e-3
    final_boxes[:, 2] -= final_boxes[:, 0]                                # drivers mutate outputs in place
    again = sess.run(model.final_boxes, feed_dict=feed_dict)
    assert np.abs(again - ref["final_boxes"]).max() < 1e-3


def test_batch_surface_like_multi_queuer():
    from object_detection_tracking_b200 import backend
    from object_detection_tracking_b200.config import make_config
    from object_detection_tracking_b200.synth import synth_frame, synth_weights
    cfg = make_config(resnet_num_block=(1, 1, 1, 1), max_size=256, short_edge_size=192, im_batch_size=3)
    model = backend.get_model(cfg, gpuid=0, is_multi=True)
    model.set_weights(synth_weights(cfg, 7))
    sess = backend.Session()
    imgs = [synth_frame(192, 256, s).astype("float32") for s in range(3)]
    fd = model.get_feed_dict_forward_multi(imgs)                          # models.py:3301-3310
    boxes, labels, probs, valid, feats = sess.run(
        [model.final_boxes, model.final_labels, model.final_probs, model.final_valid_indices, model.fpn_box_feat],
        feed_dict=fd)                                                     # multi_queuer.py:474-479
    assert boxes.shape == (3, 100, 4) and labels.dtype == np.float32 and valid.dtype == np.int32
    assert sum(valid) == feats.shape[0]                                   # multi_queuer.py:480


def test_missing_weights_fail_loudly():
    from object_detection_tracking_b200 import backend
    from object_detection_tracking_b200.config import make_config
    model = backend.get_model(make_config(resnet_num_block=(1, 1, 1, 1)))
    with pytest.raises(RuntimeError):
        backend.Session().run(model.final_boxes, feed_dict=model.get_feed_dict_forward(np.zeros((64, 64, 3), "float32")))


def test_uint8_ingest_and_pooled_features_are_the_same_detections():
    """The two boundary extensions a driver can opt into: uint8 frames across PCIe and mean-pooled
    box features (`feat_mode=1`, the np.mean of multi_queuer.py:484-485 on the GPU) -- same boxes / labels / probs as the
    float32-frame, [R,256,7,7] reference semantics on the same pixel values; the pooled features equal the host mean."""
    from object_detection_tracking_b200 import backend
    from object_detection_tracking_b200.config import make_config
    from object_detection_tracking_b200.synth import synth_frame, synth_weights
    cfg = make_config(resnet_num_block=(1, 1, 1, 1), max_size=256, short_edge_size=192)
    Wt = synth_weights(cfg, 7)
    frame_u8 = synth_frame(192, 256, 4)
    ref_model = backend.get_model(cfg, gpuid=0)
    opt_model = backend.get_model(cfg, gpuid=0, input_dtype="uint8", feat_mode=1)
    outs = []
    for model, frame in ((ref_model, frame_u8.astype(np.float32)), (opt_model, frame_u8)):
        model.set_weights(Wt)
        outs.append(backend.Session().run([model.final_boxes, model.final_labels, model.final_probs, model.fpn_box_feat],
                                          feed_dict=model.get_feed_dict_forward(frame)))
    (b0, l0, p0, f0), (b1, l1, p1, f1) = outs
    assert len(b0) > 0
    np.testing.assert_array_equal(l0, l1)
    np.testing.assert_array_equal(b0, b1)
    np.testing.assert_array_equal(p0, p1)
    assert f0.shape[1:] == (256, 7, 7) and f1.shape == (len(b1), 256)
    assert np.abs(f1 - f0.mean(axis=(2, 3))).max() <= 1e-5 * max(1.0, np.abs(f0).max())
