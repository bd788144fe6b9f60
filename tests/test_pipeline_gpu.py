"""Parity of the CUDA hot path (through the C ABI) with the CPU oracle, stage by stage and end to end.

Tolerances (written per assertion): integer outputs (counts, labels, kept sets) bit-exact; box
coordinates and scores within 1e-3 absolute in split precision (north_star); feature maps relative."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-12))


def unmatched(a, b, tol):
    """Number of rows of `a` with no row of `b` within `tol` (max-abs); a, b: [n, d].  Proposal / detection
    lists are compared as sets because scores that agree to ~1e-5 may swap the order of near-ties."""
    if len(a) == 0:
        return 0
    d = np.abs(a[:, None, :].astype(np.float64) - b[None, :, :].astype(np.float64)).max(-1)
    return int((d.min(1) > tol).sum())


@pytest.fixture(scope="module")
def small():
    from object_detection_tracking_b200.config import make_config
    from object_detection_tracking_b200.engine import Detector
    from object_detection_tracking_b200.synth import synth_frame, synth_weights
    from oracle import frcnn
    H, W, B = 192, 256, 2
    cfg = make_config(resnet_num_block=(1, 2, 2, 1), max_size=W, short_edge_size=H)
    Wt = synth_weights(cfg, 1234)
    frames = np.stack([synth_frame(H, W, seed=s) for s in (3, 4)]).astype(np.float32)
    det = Detector(cfg, B, H, W, precision="split", use_cuda_graph=False)
    det.load_weights(Wt)
    det.set_stage("image", frames)
    det.run_phases(255)
    ref = [frcnn.forward(cfg, Wt, frames[b]) for b in range(B)]
    yield cfg, Wt, frames, det, ref
    det.close()


def test_backbone_fpn_stages(small):
    cfg, Wt, frames, det, ref = small
    for b in range(2):
        for i in range(4):
            g = det.get_stage("c%d" % (i + 2))[b].transpose(2, 0, 1)
            assert rel(g, ref[b]["c2345"][i]) < 2e-5
        for i in range(5):
            r = ref[b]["p23456"][i]
            g = det.get_stage("p%d" % (i + 2))[b].transpose(2, 0, 1)[:, :r.shape[1], :r.shape[2]]
            assert rel(g, r) < 5e-5


def test_rpn_logits_and_proposals(small):
    cfg, Wt, frames, det, ref = small
    K = cfg.rpn_test_post_nms_topk
    for b in range(2):
        for i in range(5):
            g = det.get_stage("rpn_l%d" % i)[b]
            cls, box = ref[b]["rpn"][i]
            assert rel(g[..., :3], cls) < 1e-4 and rel(g[..., 3:15].reshape(box.shape), box) < 1e-4
        cnt = det.get_stage("lvl_count")[b].reshape(-1)
        lb = det.get_stage("lvl_boxes")[b].reshape(5, K, 4)
        ls = det.get_stage("lvl_scores")[b].reshape(5, K)
        for i in range(5):
            rb, rs = ref[b]["level_proposals"][i]
            assert int(cnt[i]) == len(rs)                                   # kept set size: exact
            assert np.abs(lb[i, :len(rs)] - rb).max() < 1e-3               # px
            assert np.abs(ls[i, :len(rs)] - rs).max() < 1e-4
        pc = int(det.get_stage("proposal_count")[b].reshape(-1)[0])
        assert pc == len(ref[b]["proposal_scores"])
        pb = det.get_stage("proposal_boxes")[b].reshape(K, 4)
        ps = det.get_stage("proposal_scores")[b].reshape(K)
        # same proposal set (order of near-tied scores may differ; at most the last-ranked one may flip)
        got = np.concatenate([pb[:pc], ps[:pc, None]], 1)
        exp = np.concatenate([ref[b]["proposal_boxes"], ref[b]["proposal_scores"][:, None]], 1)
        assert unmatched(got, exp, 1e-3) <= 1
        assert np.all(np.diff(ps[:pc]) <= 0)


def test_roi_head_and_final_outputs(small):
    """ROIAlign + box head + post-processing on the ORACLE's proposals (index-aligned comparison)."""
    cfg, Wt, frames, det, ref = small
    K, R, nc = cfg.rpn_test_post_nms_topk, cfg.result_per_im, cfg.num_class
    pb = np.zeros((2, K, 4, 1), np.float32); pcnt = np.zeros((2, 1, 1, 1), np.int32)
    for b in range(2):
        n = len(ref[b]["proposal_scores"])
        pb[b, :n, :, 0] = ref[b]["proposal_boxes"]
        pcnt[b] = n
    det.set_stage("proposal_boxes", pb)
    det.set_stage("proposal_count", pcnt)
    det.run_phases(16 | 32 | 64 | 128)
    for b in range(2):
        n = len(ref[b]["proposal_scores"])
        rf = det.get_stage("roi_feat").reshape(-1, 7, 7, 256)[b * K:b * K + n].transpose(0, 3, 1, 2)
        assert rel(rf, ref[b]["roi_feat"]) < 5e-5
        hl = det.get_stage("head_logits")[b, :n, :, 0]
        assert rel(hl[:, :nc], ref[b]["cls_logits"]) < 1e-4
        assert rel(hl[:, nc + 4:nc + 4 * nc].reshape(n, nc - 1, 4), ref[b]["box_logits"]) < 1e-4
        fc = int(det.get_stage("final_count")[b].reshape(-1)[0])
        assert fc == len(ref[b]["final_probs"])                             # R: exact
        fl = det.get_stage("final_labels")[b].reshape(-1)[:fc]
        np.testing.assert_array_equal(fl, ref[b]["final_labels"])          # class ids: bit-exact
        fb = det.get_stage("final_boxes")[b].reshape(R, 4)[:fc]
        fp = det.get_stage("final_probs")[b].reshape(-1)[:fc]
        assert np.abs(fb - ref[b]["final_boxes"]).max() < 1e-3              # north_star tolerance, px
        assert np.abs(fp - ref[b]["final_probs"]).max() < 1e-3
        bf = det.get_stage("fpn_box_feat")[b * R:b * R + fc]
        assert rel(bf, ref[b]["fpn_box_feat"]) < 5e-5
    det.run_phases(255)


def test_end_to_end_detections_match_oracle(small):
    cfg, Wt, frames, det, ref = small
    R = cfg.result_per_im
    det.set_stage("image", frames)
    det.run_phases(255)
    for b in range(2):
        fc = int(det.get_stage("final_count")[b].reshape(-1)[0])
        assert fc == len(ref[b]["final_probs"])
        fl = det.get_stage("final_labels")[b].reshape(-1)[:fc].astype(np.float64)
        fb = det.get_stage("final_boxes")[b].reshape(R, 4)[:fc]
        fp = det.get_stage("final_probs")[b].reshape(-1)[:fc]
        # (label, box, prob) triples as a set: labels exact (tolerance < 1), boxes/probs within 1e-3
        got = np.concatenate([fl[:, None] * 10.0, fb, fp[:, None]], 1)
        exp = np.concatenate([ref[b]["final_labels"][:, None] * 10.0, ref[b]["final_boxes"], ref[b]["final_probs"][:, None]], 1)
        assert unmatched(got, exp, 1e-3) == 0 and unmatched(exp, got, 1e-3) == 0


def test_postprocess_kernels_bit_exact_on_oracle_inputs(small):
    """Feed the oracle's own fp32 RPN logits to the proposal kernels: selected sets and order must be
    identical, coordinates equal up to expf ulps."""
    cfg, Wt, frames, det, ref = small
    K = cfg.rpn_test_post_nms_topk
    for i in range(5):
        shape, _ = det.stage_shape("rpn_l%d" % i)
        buf = np.zeros(shape, np.float32)
        for b in range(2):
            cls, box = ref[b]["rpn"][i]
            buf[b, :, :, :3] = cls
            buf[b, :, :, 3:15] = box.reshape(box.shape[0], box.shape[1], 12)
        det.set_stage("rpn_l%d" % i, buf)
    det.run_phases(8)    # proposals only
    for b in range(2):
        cnt = det.get_stage("lvl_count")[b].reshape(-1)
        lb = det.get_stage("lvl_boxes")[b].reshape(5, K, 4)
        ls = det.get_stage("lvl_scores")[b].reshape(5, K)
        for i in range(5):
            rb, rs = ref[b]["level_proposals"][i]
            assert int(cnt[i]) == len(rs)
            np.testing.assert_array_equal(ls[i, :len(rs)], rs)              # same logits selected, same order
            assert np.abs(lb[i, :len(rs)] - rb).max() < 2e-4
    det.run_phases(255)  # restore the context's own state for later tests


def test_fp16_precision_mode_is_close(small):
    from object_detection_tracking_b200.engine import Detector
    cfg, Wt, frames, det, ref = small
    d16 = Detector(cfg, 2, 192, 256, precision="fp16", use_cuda_graph=False)
    d16.load_weights(Wt)
    d16.set_stage("image", frames)
    d16.run_phases(255)
    for i in range(4):
        g = d16.get_stage("c%d" % (i + 2))[0].transpose(2, 0, 1)
        assert rel(g, ref[0]["c2345"][i]) < 2e-2          # fp16 operands: ~1e-3 per layer
    d16.close()


def test_cuda_graph_replay_equals_eager_and_outputs_are_fresh(small):
    from object_detection_tracking_b200.engine import Detector
    cfg, Wt, frames, det, ref = small
    dg = Detector(cfg, 2, 192, 256, precision="split", use_cuda_graph=True)
    dg.load_weights(Wt)
    o1 = dg.detect_host(frames)
    o2 = dg.detect_host(frames)
    for k in o1:
        np.testing.assert_array_equal(o1[k], o2[k])
        assert o1[k] is not o2[k]
    for b in range(2):
        r = int(o1["valid"][b])
        assert r == len(ref[b]["final_probs"])
        got = np.concatenate([o1["labels"][b, :r, None] * 10.0, o1["boxes"][b, :r]], 1)
        exp = np.concatenate([ref[b]["final_labels"][:, None] * 10.0, ref[b]["final_boxes"]], 1)
        assert unmatched(got, exp, 1e-3) == 0 and unmatched(exp, got, 1e-3) == 0
    # u8 frames give the same detections as the float32 copy of the same frames
    du = Detector(cfg, 2, 192, 256, precision="split", input_dtype="uint8", use_cuda_graph=True)
    du.load_weights(Wt)
    o3 = du.detect_host(frames.astype(np.uint8))
    np.testing.assert_array_equal(o3["labels"], o1["labels"])
    np.testing.assert_array_equal(o3["boxes"], o1["boxes"])
    dg.close(); du.close()


def test_pipelined_submit_wait_equals_synchronous_calls(small):
    # b2_submit_host / b2_wait (two slots, upload of batch i+1 overlapping the pass of batch i) must return exactly what
    # the synchronous b2_detect_host returns for each batch, in order
    from object_detection_tracking_b200.engine import Detector
    from object_detection_tracking_b200.synth import synth_frame
    cfg, Wt, frames, det, ref = small
    dg = Detector(cfg, 2, 192, 256, precision="split", use_cuda_graph=True)
    dg.load_weights(Wt)
    batches = [frames] + [np.stack([synth_frame(192, 256, seed=50 + 2 * j + i) for i in range(2)]).astype(np.float32)
                          for j in range(4)]
    sync = [dg.detect_host(b) for b in batches]
    outs = [dg.alloc_outputs(feat_mode=0), dg.alloc_outputs(feat_mode=0)]
    got = []
    dg.submit_host(batches[0], outs[0], 0)
    for i in range(1, len(batches)):
        dg.submit_host(batches[i], outs[i & 1], i & 1)
        dg.wait((i - 1) & 1)
        got.append({k: v.copy() for k, v in outs[(i - 1) & 1].items()})
    dg.wait((len(batches) - 1) & 1)
    got.append({k: v.copy() for k, v in outs[(len(batches) - 1) & 1].items()})
    for a, b in zip(sync, got):
        for k in a:
            np.testing.assert_array_equal(a[k], b[k])
    with pytest.raises(RuntimeError):
        dg.wait(0)                                   # nothing in flight on that slot
    dg.close()


def test_batch_graph_semantics_match_oracle_multi():
    """multi_semantics=1 (Mask_RCNN_FPN_multi: combined_non_max_suppression) vs oracle.forward_multi."""
    from object_detection_tracking_b200.config import make_config
    from object_detection_tracking_b200.engine import Detector
    from object_detection_tracking_b200.synth import synth_frame, synth_weights
    from oracle import frcnn
    H, W, B = 192, 256, 2
    cfg = make_config(resnet_num_block=(1, 1, 1, 1), max_size=W, short_edge_size=H, im_batch_size=B)
    Wt = synth_weights(cfg, 21)
    frames = np.stack([synth_frame(H, W, seed=s) for s in (8, 9)]).astype(np.float32)
    det = Detector(cfg, B, H, W, precision="split", use_cuda_graph=False, multi_semantics=True)
    det.load_weights(Wt)
    out = det.detect_host(frames)
    ref = frcnn.forward_multi(cfg, Wt, list(frames), stages=True)
    np.testing.assert_array_equal(out["valid"], ref["final_valid_indices"])
    K = cfg.rpn_test_post_nms_topk
    for b in range(B):
        # proposal set after the zero-padded merge + zero-area drop
        pc = int(det.get_stage("proposal_count")[b].reshape(-1)[0])
        assert pc == len(ref["per_image"][b]["proposal_scores"])
        pb = det.get_stage("proposal_boxes")[b].reshape(K, 4)[:pc]
        assert unmatched(pb, ref["per_image"][b]["proposal_boxes"], 1e-3) <= 1
        r = int(out["valid"][b])
        got = np.concatenate([out["labels"][b, :r, None] * 10.0, out["boxes"][b, :r], out["probs"][b, :r, None]], 1)
        exp = np.concatenate([ref["final_labels"][b, :r, None] * 10.0, ref["final_boxes"][b, :r],
                              ref["final_probs"][b, :r, None]], 1)
        assert unmatched(got, exp, 1e-3) == 0 and unmatched(exp, got, 1e-3) == 0
    # the single-image semantics differ on the same frames (score threshold / min-size / padding rules)
    single = frcnn.forward(cfg, Wt, frames[0], stages=True)
    assert len(single["proposal_scores"]) >= len(ref["per_image"][0]["proposal_scores"])
    det.close()


def test_out_of_range_activations_fail_loudly():
    """Activations are stored as fp16 (hi, lo) planes: |x| > 65504 would put inf into the hi plane and every later layer
    would silently compute on it.  conv_tc_kernel flags such a store and every host-facing call
    returns an error; the flag does not stick to the next pass."""
    from object_detection_tracking_b200.config import make_config
    from object_detection_tracking_b200.engine import Detector
    from object_detection_tracking_b200.synth import synth_frame, synth_weights
    H, W = 192, 256
    cfg = make_config(resnet_num_block=(1, 1, 1, 1), max_size=W, short_edge_size=H)
    Wt = synth_weights(cfg, 5)
    frame = synth_frame(H, W, 2).astype(np.float32)[None]
    bad = dict(Wt)
    bad["group1/block0/conv2/W"] = Wt["group1/block0/conv2/W"] * np.float32(1.5e5)    # weights still inside the fp16 range
    det = Detector(cfg, 1, H, W, precision="split", use_cuda_graph=True)
    worse = dict(Wt)
    worse["group1/block0/conv2/W"] = Wt["group1/block0/conv2/W"] * np.float32(1e9)    # the operand planes themselves overflow
    with pytest.raises(RuntimeError, match="outside the fp16-plane range"):
        det.load_weights(worse)
    det.load_weights(bad)
    with pytest.raises(RuntimeError, match="fp16-plane range"):
        det.detect_host(frame)
    outs = det.alloc_outputs()
    det.submit_host(frame, outs, 0)
    with pytest.raises(RuntimeError, match="fp16-plane range"):
        det.wait(0)
    det.load_weights(Wt)                      # same context, conditioned weights: clean again
    out = det.detect_host(frame)
    assert int(out["valid"][0]) > 0 and np.isfinite(out["boxes"]).all()
    det.close()
