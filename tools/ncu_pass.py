"""One eager pass of the hot path bracketed by cudaProfilerStart/Stop, for ncu:
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python tools/ncu_pass.py [precision] [batch]
  ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_tc \
      -c 6 -o gpurun_out/prof python tools/ncu_pass.py split 8 conv_only
A number printed under ncu is never a bench value."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from object_detection_tracking_b200.config import make_config  # noqa: E402
from object_detection_tracking_b200.engine import Detector  # noqa: E402
from object_detection_tracking_b200.synth import synth_frame, synth_weights  # noqa: E402


def main():
    precision = sys.argv[1] if len(sys.argv) > 1 else "split"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    phases = 255
    if len(sys.argv) > 3 and sys.argv[3] == "backbone":
        phases = 1
    cfg = make_config()
    det = Detector(cfg, B, 720, 1280, precision=precision, use_cuda_graph=False)
    det.load_weights(synth_weights(cfg, 1234))
    frames = np.stack([synth_frame(720, 1280, seed=i) for i in range(B)]).astype(np.float32)
    det.set_stage("image", frames)
    det.run_phases(255)            # warm-up (one-time attribute setup, L2/TLB)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    det.run_phases(phases)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("profiled pass:", det.phase_times())


if __name__ == "__main__":
    main()
