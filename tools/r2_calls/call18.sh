#!/bin/bash
# dedicated output TMA warps (warps 2, 3) instead of issuer epilogue warp + named barrier: correctness then same-box A/B
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_conv_gpu.py tests/test_pipeline_gpu.py tests/test_reid_gpu.py tests/test_tracking_gpu.py -x -q --timeout=120 2>&1 | tail -5
[ ${PIPESTATUS[0]} -eq 0 ] || { echo "STOP: tests failed / hung"; exit 1; }
timeout 400 python -m pytest tests/test_effdet_gpu.py tests/test_baseline_configs_gpu.py::test_c1_r101_720x1280_batch1_three_frames tests/test_reid_r101_mask_gpu.py -x -q --timeout=300 2>&1 | tail -4
tools/ab_run.sh new: prev: new-2: prev-2: new-3: prev-3:
