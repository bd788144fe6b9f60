"""Library baseline ("report PyTorch/cuDNN on the same B200 for each conv shape"): the main conv shapes of
the R101-FPN pass at batch 8, timed with CUDA events through torch.nn.functional.conv2d (cuDNN) in fp16 / TF32 / fp32,
channels_last.  One JSON line per (layer, dtype): ms per call and algorithmic TF/s.  Not a bench value of this repo --
context for profiles/r1_layer_table.txt.  Usage: python tools/cudnn_layer_baseline.py [batch]"""
import json
import sys

import torch
import torch.nn.functional as F

LAYERS = [  # name, H, W, Cin, Cout, k, stride, dilation, count per frame
    ("res2 1x1 64->256", 184, 320, 64, 256, 1, 1, 1, 3),
    ("res2 3x3 64->64", 184, 320, 64, 64, 3, 1, 1, 3),
    ("res3 3x3 128->128", 92, 160, 128, 128, 3, 1, 1, 4),
    ("res3 1x1 128->512", 92, 160, 128, 512, 1, 1, 1, 4),
    ("res4 3x3 256->256", 46, 80, 256, 256, 3, 1, 1, 23),
    ("res4 1x1 256->1024", 46, 80, 256, 1024, 1, 1, 1, 23),
    ("res4 1x1 1024->256", 46, 80, 1024, 256, 1, 1, 1, 23),
    ("res5 3x3 512->512 d2", 23, 40, 512, 512, 3, 1, 2, 3),
    ("fpn p2 3x3 256->256", 184, 320, 256, 256, 3, 1, 1, 1),
    ("rpn p2 3x3 256->256", 180, 320, 256, 256, 3, 1, 1, 1),
    ("fpn lateral c2 1x1 256->256", 184, 320, 256, 256, 1, 1, 1, 1),
]


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    dev = torch.device("cuda:0")
    torch.backends.cudnn.benchmark = True
    for name, H, W, ci, co, k, s, d, cnt in LAYERS:
        flops = 2.0 * B * H * W * co * ci * k * k / (s * s)
        for mode in ("fp16", "tf32", "fp32"):
            dt = torch.float16 if mode == "fp16" else torch.float32
            torch.backends.cudnn.allow_tf32 = mode == "tf32"
            x = torch.randn(B, ci, H, W, device=dev, dtype=dt).contiguous(memory_format=torch.channels_last)
            w = torch.randn(co, ci, k, k, device=dev, dtype=dt).contiguous(memory_format=torch.channels_last)
            pad = d * (k - 1) // 2
            for _ in range(5):
                F.conv2d(x, w, stride=s, padding=pad, dilation=d)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 20
            e0.record()
            for _ in range(reps):
                F.conv2d(x, w, stride=s, padding=pad, dilation=d)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            print(json.dumps(dict(layer=name, batch=B, mode=mode, ms=round(ms, 4), tflops=round(flops / ms / 1e9, 1),
                                  per_frame_count=cnt)))


if __name__ == "__main__":
    main()
