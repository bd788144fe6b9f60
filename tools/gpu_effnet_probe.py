"""GPU probe: EfficientDet pre-processing + EfficientNet backbone (+ full detect) vs the CPU oracle."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from object_detection_tracking_b200.effdet import EffdetEngine  # noqa: E402
from object_detection_tracking_b200.effdet_config import BACKBONE_OF, make_effdet_config  # noqa: E402
from object_detection_tracking_b200.synth import synth_effdet_weights, synth_efficientnet_weights  # noqa: E402
from oracle import effdet as oe  # noqa: E402
from oracle import efficientnet as on  # noqa: E402


def synth_frame(h, w, seed=3):
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, (h // 8 + 1, w // 8 + 1, 3)).astype(np.float32)
    img = np.kron(base, np.ones((8, 8, 1), np.float32))[:h, :w]
    return np.clip(img + rng.standard_normal((h, w, 3)) * 12, 0, 255).astype(np.uint8)


def run(det, H, W, fh, fw, precision, oracle=True, **over):
    bb = BACKBONE_OF[det]
    cfg = make_effdet_config(det, H, W, **over)
    Wt = dict(synth_effdet_weights(cfg))
    Wt.update(synth_efficientnet_weights(bb))
    frame = synth_frame(fh, fw)
    eng = EffdetEngine(cfg, Wt, precision=precision, backbone=bb)
    out = eng.detect(frame)
    print("== %s/%s %dx%d frame %dx%d %s launches %d dets %d scale %.6f" % (det, bb, H, W, fh, fw, precision, eng.num_launches,
                                                                           len(out["final_probs"]), out["image_scale"]))
    if oracle:
        t0 = time.time()
        img, scale = on.preprocess(frame, H, W)
        feats = on.forward(img, Wt, bb, stages=True)
        ref = oe.forward_from_features(cfg, Wt, {l: feats[l] for l in (3, 4, 5)}, image_scale=scale, stages=True)
        print("  oracle %.1fs scale %.6f" % (time.time() - t0, scale))
        g = eng.stage("image")
        print("  image max|d| %.3e" % np.abs(g - img).max())
        st = feats["stages"]
        g = eng.stage("stem", real=st["stem"].shape[0])
        print("  stem max|d| %.3e (ref max %.2f)" % (np.abs(g - st["stem"].transpose(1, 2, 0)).max(), np.abs(st["stem"]).max()))
        nb = len(st) - 1
        for i in sorted(set([0, 1, 2, nb // 2, nb - 1])):
            r = st["block_%d" % i].transpose(1, 2, 0)
            g = eng.stage("block_%d" % i, real=r.shape[2])
            print("  block_%d max|d| %.3e (ref max %.2f)" % (i, np.abs(g - r).max(), np.abs(r).max()))
        for l in (3, 4, 5):
            r = feats[l].transpose(1, 2, 0)
            print("  c%d max|d| %.3e (ref max %.2f)" % (l, np.abs(eng.stage("c%d" % l) - r).max(), np.abs(r).max()))
        for l in range(3, 8):
            r = ref["cls_out"][l]
            print("  cls%d max|d| %.3e (ref max %.2f)" % (l, np.abs(eng.stage("cls%d" % l) - r).max(), np.abs(r).max()))
        n, nr = len(out["final_probs"]), len(ref["final_probs"])
        m = min(n, nr)
        print("  detections gpu %d ref %d labels equal %s levels equal %s" % (
            n, nr, np.array_equal(out["final_labels"][:m], ref["final_labels"][:m]), np.array_equal(out["levels"][:m], ref["levels"][:m])))
        if m:
            print("  boxes max|d| %.3e px scores max|d| %.3e feat max|d| %.3e" % (
                np.abs(out["final_boxes"][:m] - ref["final_boxes"][:m]).max(), np.abs(out["final_probs"][:m] - ref["final_probs"][:m]).max(),
                np.abs(out["fpn_box_feat"][:m] - ref["fpn_box_feat"][:m]).max()))
    for _ in range(3):
        eng.detect(frame)
    t0 = time.time()
    for _ in range(10):
        eng.detect(frame)
    print("  detect() host-call time %.2f ms/frame" % ((time.time() - t0) / 10 * 1e3))
    eng.close()


if __name__ == "__main__":
    run("efficientdet-d0", 256, 384, 300, 500, "split", fpn_cell_repeats=2, box_class_repeats=2)
    run("efficientdet-d0", 256, 384, 300, 500, "fp16", fpn_cell_repeats=2, box_class_repeats=2)
    run("efficientdet-d1", 256, 256, 200, 190, "split", fpn_cell_repeats=1, box_class_repeats=1)
    if "--full" in sys.argv:
        run("efficientdet-d0", 512, 512, 720, 1280, "split", oracle=True)
        run("efficientdet-d7", 1536, 1536, 1080, 1920, "split", oracle=False)
        run("efficientdet-d7", 1536, 1536, 1080, 1920, "fp16", oracle=False)
