#!/bin/bash
# First gpurun call of a session that follows GPU-less work (end of round 1): run the not-yet-verified GPU tests in
# isolation first (each file in its own process, under its own timeout, so a fault there cannot take the verified suite
# or the box down with it), then the regular suite, then one bench line.
#   gpurun --timeout 1500 -- 'bash tools/first_gpu_call.sh'
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== widened pieces (tests/test_zz_widen_gpu.py) =="
timeout 300 python -m pytest tests/test_zz_widen_gpu.py -m gpu -q --runxfail --timeout=120 2>&1 | tee gpurun_out/widen_tests.log | tail -25
echo "== opt-in engines (tests/test_zz_unverified_gpu.py, B2_RUN_UNVERIFIED=1) =="
B2_RUN_UNVERIFIED=1 timeout 400 python -m pytest tests/test_zz_unverified_gpu.py -m gpu -q --timeout=150 2>&1 | tee gpurun_out/unverified_tests.log | tail -25
echo "== regular suite =="
timeout 600 python -m pytest tests -m gpu -x -q --timeout=150 --durations=5 2>&1 | tee gpurun_out/gpu_tests.log | tail -14
echo "== bench =="
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_first.json 2> gpurun_out/bench_first.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/bench_first.json
echo "== aux timing (JDE tracker on the GPU cost, pair cost, resize) =="
timeout 200 python tools/gpu_widen_timing.py > gpurun_out/widen_timing.jsonl 2> gpurun_out/widen_timing.err; echo "timing rc=$?"; cat gpurun_out/widen_timing.jsonl
echo "== cuDNN library baseline for the main conv shapes =="
timeout 200 python tools/cudnn_layer_baseline.py 8 > gpurun_out/cudnn_layers_b8.jsonl 2> gpurun_out/cudnn_layers.err; echo "cudnn rc=$?"; tail -5 gpurun_out/cudnn_layers_b8.jsonl
echo "== two contexts in flight (experiment) =="
timeout 300 python tools/ab_two_contexts.py 40 > gpurun_out/ab_two_contexts.jsonl 2> gpurun_out/ab_two_contexts.err; echo "rc=$?"; cat gpurun_out/ab_two_contexts.jsonl
# Two-GPU test of the peer-memory exchange (config 5) needs its own call:
#   gpurun --gpus 2 --timeout 600 -- 'B2_RUN_UNVERIFIED=1 timeout 400 python -m pytest tests/test_zz_multi_gpu.py -m gpu -q --runxfail'
