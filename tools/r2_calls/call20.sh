#!/bin/bash
# L2 -> shared-memory fill bandwidth probe (what bounds conv_tc_kernel's operand ring): tools/l2_fill_probe.cu
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 90 build/l2_fill_probe | tee gpurun_out/r2_l2_fill_probe.jsonl
echo "rc=$?"
