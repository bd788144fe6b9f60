"""GPU probe: OSNet stage-by-stage comparison against reference intermediates (debug fixture)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from object_detection_tracking_b200.reid import ReidEngine
from object_detection_tracking_b200.synth import synth_osnet_state
g = np.load(os.path.join(ROOT, "tests/golden/osnet.npz"))
st = np.load(os.path.join(ROOT, "tests/golden/_debug_osnet_stages.npz"))
eng = ReidEngine(batch=1, precision=sys.argv[1] if len(sys.argv) > 1 else "split")
eng.load_state(synth_osnet_state(4321))
for it in range(3):
  f = eng.embed(g["resized"][:1])
  print("== pass", it, "feat maxabs err", np.abs(f - st["feat"]).max(), "ref max", np.abs(st["feat"]).max())
  for k in ["conv1", "maxpool", "conv2.0.x1", "conv2.0.s0", "conv2.0.s1", "conv2.0.s3", "conv2.0", "conv2.1", "conv2.t", "conv2", "conv3.0", "conv3", "conv4.1", "conv5"]:
    a = eng.get_activation(k)
    r = st[k].astype(np.float32)
    a = a[..., :r.shape[-1]]
    e = np.abs(a - r)
    bad = np.argwhere(e > 0.05)
    print("   %-12s maxerr %.3e bad %d %s" % (k, e.max(), len(bad), ("rows%s cols%s" % (sorted(set((bad[:,1]*a.shape[2]+bad[:,2]).tolist()))[:6], sorted(set(bad[:,3].tolist()))[:6])) if len(bad) else ""))
