"""Per-kind / per-step timing table of the EfficientDet pass (GPU).  Usage: python tools/effdet_layer_report.py
[det] [H] [W] [precision] > profiles/...txt"""
import sys

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tools")
from gpu_effnet_probe import synth_frame  # noqa: E402
from object_detection_tracking_b200.effdet import EffdetEngine  # noqa: E402
from object_detection_tracking_b200.effdet_config import BACKBONE_OF, make_effdet_config  # noqa: E402
from object_detection_tracking_b200.synth import synth_effdet_weights, synth_efficientnet_weights  # noqa: E402

det = sys.argv[1] if len(sys.argv) > 1 else "efficientdet-d7"
H = int(sys.argv[2]) if len(sys.argv) > 2 else 1536
W = int(sys.argv[3]) if len(sys.argv) > 3 else 1536
prec = sys.argv[4] if len(sys.argv) > 4 else "split"
cfg = make_effdet_config(det, H, W)
Wt = dict(synth_effdet_weights(cfg))
Wt.update(synth_efficientnet_weights(BACKBONE_OF[det]))
eng = EffdetEngine(cfg, Wt, precision=prec, backbone=BACKBONE_OF[det])
frame = synth_frame(1080, 1920)
eng.detect(frame)
steps = eng.profile_steps(5)
kinds = ["pointwise conv (tcgen05)", "depthwise 3x3 (BiFPN/heads)", "combine / pool", "depthwise+BN+swish (backbone)",
         "squeeze-excite", "stem im2col", "post-processing", "box feature"]
tot = sum(s[2] for s in steps)
print("# %s %dx%d %s: %d steps, %.3f ms eager sum" % (det, H, W, prec, len(steps), tot))
print("kind,steps,ms,share,GB/s,TFLOP/s")
for k, nm in enumerate(kinds):
    sel = [s for s in steps if s[1] == k]
    if not sel:
        continue
    ms = sum(s[2] for s in sel)
    print("%s,%d,%.3f,%.1f%%,%.0f,%.1f" % (nm, len(sel), ms, 100 * ms / tot, sum(s[4] for s in sel) / ms / 1e6,
                                          sum(s[3] for s in sel) / ms / 1e9))
print("\n# slowest 40 steps\nname,kind,ms,GB/s,TFLOP/s")
for s in sorted(steps, key=lambda t: -t[2])[:40]:
    print("%s,%d,%.4f,%.0f,%.1f" % (s[0], s[1], s[2], s[4] / s[2] / 1e6, s[3] / s[2] / 1e9))
if len(sys.argv) > 5:
    with open(sys.argv[5], "w") as f:
        f.write("idx,name,kind,ms,flops,bytes\n")
        for i, st in enumerate(steps):
            f.write("%d,%s,%d,%.5f,%.0f,%.0f\n" % (i, st[0].replace(",", ";"), st[1], st[2], st[3], st[4]))
import time
for _ in range(3):
    eng.detect(frame)
t0 = time.time()
for _ in range(10):
    eng.detect(frame)
print("\n# detect() host call (graph replay, incl. H2D frame + D2H results): %.2f ms/frame" % ((time.time() - t0) / 10 * 1e3))
