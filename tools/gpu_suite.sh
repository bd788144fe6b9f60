#!/bin/bash
# The whole -m gpu suite + one bench line on a B200:   gpurun --timeout 1500 -- 'bash tools/gpu_suite.sh'
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/baseline_parity.jsonl
timeout 1200 python -m pytest tests -m gpu -x -q --timeout=400 --durations=8 2>&1 | tee gpurun_out/gpu_tests.log | tail -20
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 1800 gpurun_out/bench.json
