#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/baseline_parity.jsonl
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 --durations=6 2>&1 | tail -40
echo "=== margins"; cut -c1-700 gpurun_out/baseline_parity.jsonl
echo "=== conv + pipeline parity under B2_BK32=1"
B2_BK32=1 timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_pipeline_gpu.py tests/test_reid_gpu.py -x -q 2>&1 | tail -5
echo "=== conv + pipeline parity under B2_L2_PREFETCH=2"
B2_L2_PREFETCH=2 timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_pipeline_gpu.py -x -q 2>&1 | tail -5
tools/ab_run.sh base: bk32:B2_BK32=1 bk32res:B2_BK32=2 l2pf:B2_L2_PREFETCH=1 l2pf2:B2_L2_PREFETCH=2 bk32l2pf:B2_BK32=1,B2_L2_PREFETCH=1 pdl-bk32:B2_BK32=1 base2:
