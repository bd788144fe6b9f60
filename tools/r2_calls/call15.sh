#!/bin/bash
# PDL (programmatic dependent launch) A/B with the final lean kernel
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
P=object_detection_tracking_b200
cp $P/libb200det.so /tmp/main.so; cp $P/libb200det_pdl.so $P/libb200det.so
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_pipeline_gpu.py tests/test_baseline_configs_gpu.py::test_c1_r101_720x1280_batch1_three_frames tests/test_reid_gpu.py -x -q --timeout=300 2>&1 | tail -4
cp /tmp/main.so $P/libb200det.so
tools/ab_run.sh base: pdl: base2: pdl2: base3: pdl3:
