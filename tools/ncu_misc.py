"""The small kernels added at the end of round 1, each launched once between cudaProfilerStart/Stop, for ncu:
resize_u8_to_f32 (ingest), pair_segmin (multi-camera pair cost), mask_select + roialign_kernel<14> (--add_mask),
agg_feat (feat_mode 2/3 box-feature aggregation)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from object_detection_tracking_b200 import engine, reid  # noqa: E402
from object_detection_tracking_b200.config import make_config  # noqa: E402
from object_detection_tracking_b200.synth import synth_frame, synth_weights  # noqa: E402

rng = np.random.default_rng(0)
src = rng.integers(0, 256, (8, 1080, 1920, 3)).astype(np.uint8)
H, W = 384, 640
cfg = make_config(resnet_num_block=(1, 1, 2, 1), max_size=W, short_edge_size=H, add_mask=True)
det = engine.Detector(cfg, 2, H, W, precision="split", use_cuda_graph=False)
det.load_weights(synth_weights(cfg, 1234))
frames = np.stack([synth_frame(H, W, seed=s) for s in (1, 2)]).astype(np.float32)
N, M, per, D = 50, 50, 60, 512
a = rng.standard_normal((N * per, D)).astype(np.float32)
b = rng.standard_normal((M * per, D)).astype(np.float32)
sa = (np.arange(N + 1) * per).astype(np.int32)
sb = (np.arange(M + 1) * per).astype(np.int32)
tr1 = {i: (np.array([[0, 0.0, 0.0]]), a[sa[i]:sa[i + 1]]) for i in range(N)}
tr2 = {i: (np.array([[0, 0.0, 0.0]]), b[sb[i]:sb[i + 1]]) for i in range(M)}


def work():
    engine.resize_frames(src, 1280, 720)
    det.detect_host(frames)                       # roialign<14> + mask head + mask_select
    det.get_masks()
    det.detect_host(frames, feat_mode=2)          # agg_feat (max over the 7x7 grid)
    det.detect_host(frames, feat_mode=3)          # agg_feat (spatial)
    reid.compute_feature_dist(tr1, tr2, np.zeros((N, M)))


work()
torch.cuda.synchronize()
torch.cuda.profiler.start()
work()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
