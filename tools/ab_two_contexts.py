"""Round-2 experiment (not a bench value): does keeping TWO batches in flight pay?  The serial tail of a pass (proposals,
ROIAlign, head, post-processing: ~1.3 ms of 16, at most 40 blocks wide) leaves most SMs idle; with two contexts on two
streams the other batch's backbone kernels fill them.  Prints device-resident FPS with 1 and with 2 alternating contexts
(same weights, CUDA-graph replay, 4 rotating input batches each), wall clock between device synchronisations.
Usage: python tools/ab_two_contexts.py [steps]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from object_detection_tracking_b200.config import make_config
    from object_detection_tracking_b200.engine import Detector
    from object_detection_tracking_b200.synth import synth_frame, synth_weights
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    B, H, W = 8, 720, 1280
    cfg = make_config()
    Wt = synth_weights(cfg, 1234)
    dets = []
    for k in range(2):
        d = Detector(cfg, B, H, W, device=0, precision="split", use_cuda_graph=True)
        d.load_weights(Wt)
        dets.append(d)
    dev = [torch.from_numpy(np.stack([synth_frame(H, W, seed=8 * j + i) for i in range(B)]).astype(np.float32)).cuda()
           for j in range(4)]
    ref = None
    for n_ctx in (1, 2, 1, 2):
        use = dets[:n_ctx]
        for i in range(4):
            use[i % n_ctx].detect_device(dev[i % 4], None, sync=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(steps):
            use[i % n_ctx].detect_device(dev[i % 4], None, sync=False)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        out = use[(steps - 1) % n_ctx].get_stage("final_boxes")
        if ref is None:
            ref = out.copy()
        print(json.dumps(dict(contexts=n_ctx, steps=steps, fps=steps * B / dt, ms_per_step=dt / steps * 1e3,
                              same_boxes_as_first_run=bool(np.array_equal(out, ref)))), flush=True)


if __name__ == "__main__":
    main()
