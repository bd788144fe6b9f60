#!/bin/bash
# CTA pairs (cta_group::2) bring-up: bit-identity vs the single-CTA path on the diag build, then the conv / pipeline tests
# with pairs on, then a same-box A/B of the bench (pairs on the layers with >= 8 K-blocks)
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
P=object_detection_tracking_b200
cp $P/libb200det.so /tmp/main.so
cp $P/libb200det_pairdiag.so $P/libb200det.so
timeout -k 5 150 python tools/pair_check.py 1 > gpurun_out/pair_check.log 2>&1; rc=$?
cp /tmp/main.so $P/libb200det.so
tail -45 gpurun_out/pair_check.log
[ $rc -eq 0 ] || { echo "STOP: pair_check rc=$rc"; exit 1; }
B2_PAIR=1 timeout -k 5 150 python -m pytest tests/test_conv_gpu.py -x -q --timeout=60 2>&1 | tail -4
[ ${PIPESTATUS[0]} -eq 0 ] || { echo "STOP: conv tests with pairs"; exit 1; }
B2_PAIR=8 timeout -k 5 200 python -m pytest tests/test_pipeline_gpu.py tests/test_baseline_configs_gpu.py::test_c1_r101_720x1280_batch1_three_frames -x -q --timeout=120 2>&1 | tail -4
[ ${PIPESTATUS[0]} -eq 0 ] || { echo "STOP: pipeline tests with pairs"; exit 1; }
for v in base pair base-2 pair-2; do
  if [ "${v%%-*}" = "pair" ]; then export B2_PAIR=8; else unset B2_PAIR; fi
  timeout -k 5 100 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-stream --sustained-seconds 0 \
     --profile-json gpurun_out/layers_$v.json > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
  echo "== $v rc=$?"; python -c "
import json,sys
try:
    d=json.loads(open('gpurun_out/bench_$v.json').read().strip().splitlines()[-1]); print('   %.1f FPS  %.3f ms/step  e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']))
except Exception as e: print('   no bench line', e)
"; tail -2 gpurun_out/bench_$v.err
done
