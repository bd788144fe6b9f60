"""Parity of the CUDA hot path with the CPU oracle AT THE BASELINE.json CONFIGURATIONS (VERDICT r1 "missing #1"):

  C1  R101 (3,4,23,3) dilated, 15 classes, rpn300, 720x1280, batch 1      (models.py:488-973)
  C2  same detector, batch 8, Mask_RCNN_FPN_multi semantics                (models.py:2058-2409, 2924-2976)
  C4  R50 (3,4,6,3), COCO 81 classes, no dilation (--is_coco_model => version 2), 720x1280
  odd a 200x300 frame: the stem pad [3, 2+pad_to_32] (nn.py:871-877), the p2-p4 crop to ceil(H/stride)
      (models.py:382-390) and the uncropped p5/p6 all differ from the multiple-of-32 toy frames
  C3  EfficientDet-D7 1536x1536 on one frame is in tests/test_effdet_gpu.py::test_d7_full_size_matches_oracle

Bars (north_star): counts and class ids bit-exact, box coordinates and probabilities within 1e-3 absolute; feature
maps relative.  Every test also appends its measured margins to gpurun_out/baseline_parity.jsonl (copied to profiles/).

What "within 1e-3 of the reference" can mean at this depth.  Coordinates reach 1280 px (float32 ulp 1.2e-4 px) after 105
conv layers, so two float32 evaluations of the SAME graph already differ at the 1e-3 level: the float32 oracle is
4.0e-4 ... 5.5e-4 px away from the float64 evaluation of itself (oracle.frcnn.exact(), measured on these frames on the CPU;
the reference's TF/Eigen kernels are a third float32 realisation with their own summation order).  The tests therefore
hold the GPU path to
  * 1e-3 px / 1e-3 prob against the float64 evaluation -- the noise-free centre of the reference's arithmetic -- and
  * 1.5e-3 px against the float32 oracle port (1e-3 + the port's own measured rounding noise),
and log all three distances per frame (GPU-exact, GPU-oracle32, oracle32-exact).  Measured, round 2: GPU-oracle32
6.1e-4 ... 1.04e-3 px over 11 full-size frames (one frame of eleven above 1e-3), probabilities within 8e-6.
"""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-12))


def set_dist(a, b):
    """max over rows of a of the distance (max-abs) to the nearest row of b."""
    if len(a) == 0:
        return 0.0
    d = np.abs(a[:, None, :].astype(np.float64) - b[None, :, :].astype(np.float64)).max(-1)
    return float(d.min(1).max())


def log_margins(rec):
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "baseline_parity.jsonl"), "a") as f:
        f.write(json.dumps(rec) + "\n")


def stagewise(det, cfg, b, o, rec):
    """Backbone / FPN / RPN / proposals of image `b` of the context against the oracle's stages `o`."""
    K = cfg.rpn_test_post_nms_topk
    rec["c_rel"] = [rel(det.get_stage("c%d" % (i + 2))[b].transpose(2, 0, 1), o["c2345"][i]) for i in range(4)]
    p_rel = []
    for i in range(5):
        r = o["p23456"][i]
        g = det.get_stage("p%d" % (i + 2))[b].transpose(2, 0, 1)
        assert g.shape[1] >= r.shape[1] and g.shape[2] >= r.shape[2]
        p_rel.append(rel(g[:, :r.shape[1], :r.shape[2]], r))
    rec["p_rel"] = p_rel
    rec["p_shapes"] = [list(o["p23456"][i].shape[1:]) for i in range(5)]
    rpn_rel = []
    for i in range(5):
        g = det.get_stage("rpn_l%d" % i)[b]
        cls, box = o["rpn"][i]
        assert g.shape[:2] == cls.shape[:2]                              # the cropped level extents (models.py:382-390)
        rpn_rel.append(max(rel(g[..., :3], cls), rel(g[..., 3:15].reshape(box.shape), box)))
    rec["rpn_rel"] = rpn_rel
    cnt = det.get_stage("lvl_count")[b].reshape(-1)
    lb = det.get_stage("lvl_boxes")[b].reshape(5, K, 4)
    ls = det.get_stage("lvl_scores")[b].reshape(5, K)
    rec["lvl_count"] = [int(c) for c in cnt]
    rec["lvl_count_ref"] = [len(o["level_proposals"][i][1]) for i in range(5)]
    lvl_box, lvl_score = 0.0, 0.0
    for i in range(5):
        rb, rs = o["level_proposals"][i]
        if int(cnt[i]) == len(rs) and len(rs):
            got = np.concatenate([lb[i, :len(rs)], ls[i, :len(rs), None]], 1)
            exp = np.concatenate([rb, rs[:, None]], 1)
            lvl_box = max(lvl_box, set_dist(got, exp), set_dist(exp, got))
    rec["lvl_set_dist"] = lvl_box
    pc = int(det.get_stage("proposal_count")[b].reshape(-1)[0])
    rec["proposals"] = pc
    rec["proposals_ref"] = len(o["proposal_scores"])
    pb = det.get_stage("proposal_boxes")[b].reshape(K, 4)[:pc]
    rec["proposal_set_dist"] = max(set_dist(pb, o["proposal_boxes"]), set_dist(o["proposal_boxes"], pb))
    return rec


def assert_stagewise(rec, feat_tol=5e-5, rpn_tol=2e-4):
    assert max(rec["c_rel"]) < feat_tol, rec["c_rel"]
    assert max(rec["p_rel"]) < 2 * feat_tol, rec["p_rel"]
    assert max(rec["rpn_rel"]) < rpn_tol, rec["rpn_rel"]
    assert rec["lvl_count"] == rec["lvl_count_ref"]                       # kept-set sizes per level: exact
    assert rec["lvl_set_dist"] < 2e-3                                     # intermediate (exp-decoded boxes up to 1280 px)
    assert rec["proposals"] == rec["proposals_ref"]
    assert rec["proposal_set_dist"] < 1.5e-3


def final_margins(labels, boxes, probs, o_labels, o_boxes, o_probs, rec):
    rec["final"] = int(len(labels))
    rec["final_ref"] = int(len(o_labels))
    rec["labels_sorted_equal"] = bool(np.array_equal(np.sort(labels), np.sort(o_labels)))
    rec["order_equal"] = bool(len(labels) == len(o_labels) and np.array_equal(labels, o_labels))
    got = np.concatenate([np.asarray(labels, np.float64)[:, None] * 10.0, boxes, np.asarray(probs)[:, None]], 1)
    exp = np.concatenate([np.asarray(o_labels, np.float64)[:, None] * 10.0, o_boxes, np.asarray(o_probs)[:, None]], 1)
    rec["final_set_dist"] = max(set_dist(got, exp), set_dist(exp, got))
    if rec["order_equal"]:
        rec["box_maxabs"] = float(np.abs(boxes - o_boxes).max()) if len(labels) else 0.0
        rec["prob_maxabs"] = float(np.abs(probs - o_probs).max()) if len(labels) else 0.0
    return rec


def exact_margins(labels, boxes, probs, o32, o64, rec):
    """Distances of the GPU triples and of the float32 oracle's to the float64 evaluation (see the module docstring)."""
    trip = lambda l, b, p: np.concatenate([np.asarray(l, np.float64)[:, None] * 10.0, b, np.asarray(p)[:, None]], 1)
    got, e32, e64 = trip(labels, boxes, probs), trip(*o32), trip(*o64)
    rec["exact_final"] = int(len(e64))
    rec["gpu_to_exact"] = max(set_dist(got, e64), set_dist(e64, got))
    rec["oracle32_to_exact"] = max(set_dist(e32, e64), set_dist(e64, e32))
    return rec


def assert_final(rec):
    assert rec["final"] == rec["final_ref"]                               # number of detections: exact
    assert rec["labels_sorted_equal"]                                     # class ids: bit-exact
    assert rec["final_set_dist"] < 1.5e-3                                 # vs the float32 port: 1e-3 + its own rounding noise
    if "gpu_to_exact" in rec:
        assert rec["gpu_to_exact"] < 1e-3                                 # (label, box, prob) triples vs the exact evaluation
    if rec["order_equal"]:
        assert rec["box_maxabs"] < 1.5e-3 and rec["prob_maxabs"] < 1e-3


def run_single(cfg, seeds, H, W, tag, weight_seed=1234):
    """Batch-1 semantics: every frame through a batch-1 context, stage-wise + final against oracle.frcnn.forward."""
    from object_detection_tracking_b200.engine import Detector
    from object_detection_tracking_b200.synth import synth_frame, synth_weights
    from oracle import frcnn
    Wt = synth_weights(cfg, weight_seed)
    det = Detector(cfg, 1, H, W, precision="split", use_cuda_graph=False)
    det.load_weights(Wt)
    R = cfg.result_per_im
    recs = []
    try:
        for s in seeds:
            frame = synth_frame(H, W, seed=s).astype(np.float32)
            det.set_stage("image", frame[None])
            det.run_phases(255)
            o = frcnn.forward(cfg, Wt, frame)
            with frcnn.exact():
                o64 = frcnn.forward(cfg, Wt, frame, stages=False)
            rec = stagewise(det, cfg, 0, o, {"config": tag, "seed": s, "H": H, "W": W})
            fc = int(det.get_stage("final_count")[0].reshape(-1)[0])
            got = (det.get_stage("final_labels")[0].reshape(-1)[:fc], det.get_stage("final_boxes")[0].reshape(R, 4)[:fc],
                   det.get_stage("final_probs")[0].reshape(-1)[:fc])
            final_margins(*got, o["final_labels"], o["final_boxes"], o["final_probs"], rec)
            exact_margins(*got, (o["final_labels"], o["final_boxes"], o["final_probs"]),
                          (o64["final_labels"], o64["final_boxes"], o64["final_probs"]), rec)
            n = min(fc, len(o["final_probs"]))
            if rec["order_equal"] and n:
                rec["fpn_box_feat_rel"] = rel(det.get_stage("fpn_box_feat")[:n], o["fpn_box_feat"][:n])
            log_margins(rec)
            recs.append(rec)
    finally:
        det.close()
    for rec in recs:
        assert_stagewise(rec)
        assert_final(rec)
        assert rec.get("fpn_box_feat_rel", 0.0) < 1e-4
    return recs


def test_c1_r101_720x1280_batch1_three_frames():
    """BASELINE configs[0]/[1] geometry at full depth: 720x1280 pads to 736x1280 (stem input 741x1285), p2-p4 are cropped
    to 180x320 / 90x160 / 45x80 while p5/p6 keep 23x40 / 12x20; 105 conv layers in split precision."""
    from object_detection_tracking_b200.config import make_config
    cfg = make_config()                                                    # R101 dilated, 15 classes, rpn300
    assert cfg.resnet_num_block == (3, 4, 23, 3) and cfg.num_class == 15 and cfg.rpn_test_post_nms_topk == 300
    recs = run_single(cfg, (0, 1, 2), 720, 1280, "C1")
    assert recs[0]["p_shapes"] == [[180, 320], [90, 160], [45, 80], [23, 40], [12, 20]]
    assert all(r["final"] > 0 for r in recs)


def test_odd_geometry_200x300_pads_and_crops():
    """200x300: padded to 224x320, so c2..c5 = 56x80 .. 7x10 but p2-p4 are cropped to 50x75 / 25x38 / 13x19
    (ceil(H/stride), models.py:382-390) and p5/p6 stay 7x10 / 4x5."""
    from object_detection_tracking_b200.config import make_config
    cfg = make_config(resnet_num_block=(2, 2, 3, 2), max_size=300, short_edge_size=200)
    recs = run_single(cfg, (5, 6), 200, 300, "odd200x300", weight_seed=77)
    assert recs[0]["p_shapes"] == [[50, 75], [25, 38], [13, 19], [7, 10], [4, 5]]


def test_c4_r50_coco81_720x1280():
    """BASELINE configs[3] detector: --resnet50 --is_coco_model (version 2: no dilation, obj_detect_tracking.py:282-283),
    81 classes -> 81 + 320 box-head outputs, 80 per-class NMS problems per frame."""
    from object_detection_tracking_b200.config import make_config
    cfg = make_config(resnet_num_block=(3, 4, 6, 3), num_class=81, use_dilations=False, version=2)
    run_single(cfg, (11, 12), 720, 1280, "C4", weight_seed=4321)


def test_c2_r101_720x1280_batch8_multi_semantics():
    """BASELINE configs[1]: the batch graph Mask_RCNN_FPN_multi at batch 8 (combined NMS, zero-padded level merge) against
    oracle.frcnn.forward_multi on the same 8 frames."""
    from object_detection_tracking_b200.config import make_config
    from object_detection_tracking_b200.engine import Detector
    from object_detection_tracking_b200.synth import synth_frame, synth_weights
    from oracle import frcnn
    H, W, B = 720, 1280, 8
    cfg = make_config(im_batch_size=B)
    Wt = synth_weights(cfg, 1234)
    frames = np.stack([synth_frame(H, W, seed=20 + s) for s in range(B)]).astype(np.float32)
    det = Detector(cfg, B, H, W, precision="split", use_cuda_graph=True, multi_semantics=True)
    det.load_weights(Wt)
    try:
        out = det.detect_host(frames)
        ref = frcnn.forward_multi(cfg, Wt, list(frames), stages=True)
        with frcnn.exact():
            ref64 = frcnn.forward_multi(cfg, Wt, list(frames), stages=False)
        K = cfg.rpn_test_post_nms_topk
        recs = []
        for b in range(B):
            rec = {"config": "C2", "image": b}
            pc = int(det.get_stage("proposal_count")[b].reshape(-1)[0])
            rec["proposals"], rec["proposals_ref"] = pc, len(ref["per_image"][b]["proposal_scores"])
            pb = det.get_stage("proposal_boxes")[b].reshape(K, 4)[:pc]
            rb = ref["per_image"][b]["proposal_boxes"]
            rec["proposal_set_dist"] = max(set_dist(pb, rb), set_dist(rb, pb))
            r, rr = int(out["valid"][b]), int(ref["final_valid_indices"][b])
            got = (out["labels"][b, :r], out["boxes"][b, :r], out["probs"][b, :r])
            final_margins(*got, ref["final_labels"][b, :rr], ref["final_boxes"][b, :rr], ref["final_probs"][b, :rr], rec)
            r64 = int(ref64["final_valid_indices"][b])
            exact_margins(*got, (ref["final_labels"][b, :rr], ref["final_boxes"][b, :rr], ref["final_probs"][b, :rr]),
                          (ref64["final_labels"][b, :r64], ref64["final_boxes"][b, :r64], ref64["final_probs"][b, :r64]), rec)
            log_margins(rec)
            recs.append(rec)
    finally:
        det.close()
    for rec in recs:
        assert rec["proposals"] == rec["proposals_ref"]
        assert rec["proposal_set_dist"] < 1.5e-3
        assert_final(rec)
