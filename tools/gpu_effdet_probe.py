"""GPU probe: EfficientDet BiFPN + heads + post-processing vs the CPU oracle, stage by stage."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from object_detection_tracking_b200.effdet import EffdetEngine  # noqa: E402
from object_detection_tracking_b200.effdet_config import feat_sizes, make_effdet_config  # noqa: E402
from object_detection_tracking_b200.synth import synth_effdet_weights  # noqa: E402
from oracle import effdet as oe  # noqa: E402


def synth_features(cfg, seed=5):
    rng = np.random.default_rng(seed)
    fs = feat_sizes(cfg)
    return {l: np.abs(rng.standard_normal((cfg.backbone_channels[l - 3],) + fs[l])).astype(np.float32) for l in (3, 4, 5)}


def run(name, h, w, precision, **over):
    cfg = make_effdet_config(name, h, w, **over)
    W = synth_effdet_weights(cfg)
    feats = synth_features(cfg)
    t0 = time.time()
    ref = oe.forward_from_features(cfg, W, feats, image_scale=1.25, stages=True)
    t_ref = time.time() - t0
    eng = EffdetEngine(cfg, W, precision=precision)
    out = eng.run_features(feats, image_scale=1.25)
    print("== %s %dx%d %s  oracle %.1fs  launches %d" % (name, h, w, precision, t_ref, eng.num_launches))
    for l in range(3, 8):
        g = eng.stage("fpn%d" % l)
        r = ref["fpn"][l].transpose(1, 2, 0)
        print("  fpn%d max|d| %.3e (ref max %.3f)" % (l, np.abs(g - r).max(), np.abs(r).max()))
    for l in range(3, 8):
        for k, key in (("cls", "cls_out"), ("box", "box_out")):
            g = eng.stage("%s%d" % (k, l))
            r = ref[key][l]
            print("  %s%d max|d| %.3e (ref max %.3f)" % (k, l, np.abs(g - r).max(), np.abs(r).max()))
    n, nr = len(out["final_probs"]), len(ref["final_probs"])
    print("  detections gpu %d ref %d" % (n, nr))
    m = min(n, nr)
    if m:
        print("  labels equal %s levels equal %s" % (np.array_equal(out["final_labels"][:m], ref["final_labels"][:m]),
                                                      np.array_equal(out["levels"][:m], ref["levels"][:m])))
        print("  boxes max|d| %.3e px  scores max|d| %.3e  box_feat max|d| %.3e (ref max %.3f)" % (
            np.abs(out["final_boxes"][:m] - ref["final_boxes"][:m]).max(),
            np.abs(out["final_probs"][:m] - ref["final_probs"][:m]).max(),
            np.abs(out["fpn_box_feat"][:m] - ref["fpn_box_feat"][:m]).max(), np.abs(ref["fpn_box_feat"]).max()))
    t0 = time.time()
    for _ in range(5):
        eng.run_features(feats, image_scale=1.25)
    print("  host-call time %.2f ms" % ((time.time() - t0) / 5 * 1e3))
    eng.close()


if __name__ == "__main__":
    run("efficientdet-d0", 256, 384, "split", fpn_cell_repeats=2, box_class_repeats=2)
    run("efficientdet-d1", 256, 256, "split", fpn_cell_repeats=1, box_class_repeats=1, fpn_weight_method="sum")
    run("efficientdet-d0", 256, 384, "fp16", fpn_cell_repeats=2, box_class_repeats=2)
    if "--full" in sys.argv:
        run("efficientdet-d0", 512, 512, "split")
