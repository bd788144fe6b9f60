#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== ACC truncation compensation sweep (720x1280 R101, 2 frames each)"
for d in 0 1e-8 2e-8 3e-8 5e-8 8e-8; do
  echo "--- B2_ACC_DELTA=$d"
  B2_ACC_DELTA=$d timeout 300 python tools/gpu_pipeline_probe.py 720 1280 tcgen05 split 3,4,23,3 2 2>&1 | grep EXACT
done | tee gpurun_out/acc_delta_sweep.txt
echo "=== same with two K-blocks per chunk everywhere (B2_ACC_KB=2)"
for d in 0 3e-8 6e-8 1e-7; do
  echo "--- B2_ACC_KB=2 B2_ACC_DELTA=$d"
  B2_ACC_KB=2 B2_ACC_DELTA=$d timeout 300 python tools/gpu_pipeline_probe.py 720 1280 tcgen05 split 3,4,23,3 1 2>&1 | grep EXACT
done | tee -a gpurun_out/acc_delta_sweep.txt
echo "=== bench (new fields) "
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_new.json 2> gpurun_out/bench_new.err; echo "rc=$?"; tail -3 gpurun_out/bench_new.err; python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_new.json"))
for k in ("value","ms_per_step","e2e","sustained","stream_c1","clocks","cpu_baseline"): print(k, d.get(k))
PY
tools/ab_run.sh kb2:B2_ACC_KB=2
