import os, sys
import numpy as np, torch, torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from object_detection_tracking_b200 import engine
rng = np.random.default_rng(0)
for (B, H, W, Cin, Cout) in [(1, 64, 32, 256, 256), (1, 64, 32, 256, 64), (1, 32, 16, 384, 384), (1, 16, 8, 512, 512), (1, 64, 32, 64, 256), (2, 64, 32, 256, 256), (8, 64, 32, 256, 256)]:
    x = rng.standard_normal((B, H, W, Cin)).astype(np.float32)
    w = (rng.standard_normal((1, 1, Cin, Cout)) / np.sqrt(Cin)).astype(np.float32)
    b = rng.standard_normal(Cout).astype(np.float32)
    ref = torch.relu(F.conv2d(torch.from_numpy(x).permute(0, 3, 1, 2), torch.from_numpy(w).permute(3, 2, 0, 1).contiguous(), torch.from_numpy(b))).permute(0, 2, 3, 1).numpy()
    for split in (True, False):
        out = engine.op_conv2d(x, w, b, None, relu=True, split=split)
        e = np.abs(out - ref)
        bad = np.argwhere(e > 1e-2 * np.abs(ref).max())
        print((B, H, W, Cin, Cout), "split" if split else "fp16", "rel err %.3e" % (e.max() / np.abs(ref).max()), "bad", len(bad), "first", bad[:3].tolist(), "badcols", sorted(set(bad[:, 3].tolist()))[:8] if len(bad) else [])
