#!/bin/bash
# Round-2 A/B list prepared at the end of round 1 (no GPU then): each line is one variant of tools/ab_run.sh, i.e. one
# bench run (~40 s) on the same box.  Accuracy of the variants that change arithmetic is checked with the pipeline
# probe at 720x1280 (boxes must stay within 1e-3 px of the oracle).
#   bash tools/build_variant.sh pdl -DB2_PDL=1          # here, before the call: the variant .so travels with the snapshot
#   gpurun --timeout 1500 -- 'bash tools/round2_experiments.sh'
set -u
cd "$(dirname "$0")/.."
tools/ab_run.sh \
  base: \
  shortk2:B2_ACC_KB_SHORTK=2 \
  bn64tail:B2_BN64_TAIL=1 \
  bn64tail_serial:B2_BN64_TAIL=1,B2_NO_DUAL=1 \
  serial:B2_NO_DUAL=1 \
  shortk2_bn64tail:B2_ACC_KB_SHORTK=2,B2_BN64_TAIL=1 \
  pdl: \
  pdl-serial:B2_NO_DUAL=1 \
  pdl-shortk2:B2_ACC_KB_SHORTK=2
echo "=== accuracy of the short-K chunk variant (boxes vs the fp32 oracle on the 720x1280 frame)"
B2_ACC_KB_SHORTK=2 timeout 600 python tools/gpu_pipeline_probe.py 720 1280 tcgen05 split > gpurun_out/pipe_shortk2.log 2>&1
grep -E "c[45] rel|proposals gpu|final gpu" gpurun_out/pipe_shortk2.log
echo "=== trackers with the persistent distance workspace (B2_WS=1) vs per-call allocation"
timeout 200 python tools/gpu_tracker_probe.py > gpurun_out/tracker_nows.log 2>&1; tail -2 gpurun_out/tracker_nows.log
B2_WS=1 timeout 200 python tools/gpu_tracker_probe.py > gpurun_out/tracker_ws.log 2>&1; tail -2 gpurun_out/tracker_ws.log
timeout 200 python tools/gpu_widen_timing.py | head -1; B2_WS=1 timeout 200 python tools/gpu_widen_timing.py | head -1
