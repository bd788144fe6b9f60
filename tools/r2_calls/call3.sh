#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/baseline_parity.jsonl
timeout 1000 python -m pytest tests/test_baseline_configs_gpu.py tests/test_effdet_gpu.py::test_d7_full_size_matches_oracle tests/test_widen_gpu.py::test_distance_calls_from_the_persistent_workspace_equal_the_per_call_path tests/test_tracking_gpu.py -q --timeout=600 2>&1 | tail -40
echo "=== margins"; cat gpurun_out/baseline_parity.jsonl
P=object_detection_tracking_b200
echo "=== variant all3 through the conv + pipeline parity tests"
cp $P/libb200det.so /tmp/main.so; cp $P/libb200det_all3.so $P/libb200det.so
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_pipeline_gpu.py tests/test_effdet_gpu.py -x -q --deselect tests/test_effdet_gpu.py::test_d7_full_size_matches_oracle 2>&1 | tail -5
cp /tmp/main.so $P/libb200det.so
tools/ab_run.sh base: roll: rollbias: rollpdl: all3: base2:
