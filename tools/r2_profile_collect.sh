#!/bin/bash
# Copies what a tools/r2_profile.sh call left in gpurun_out/ into the committed summaries under profiles/.
set -u
cd "$(dirname "$0")/.."
for n in r2_conv_tc_res4_block1_split r2_post_kernels_split r2_bifpn_cell r2_effnet_mbconv r2_osnet_osblock r2_misc_kernels; do
  [ -f gpurun_out/profiles_r2/${n}_raw.csv ] || { echo "missing $n"; continue; }
  cp gpurun_out/profiles_r2/${n}_raw.csv profiles/${n}_raw.csv
  echo "== $n"; python tools/ncu_raw_table.py profiles/${n}_raw.csv
done > profiles/r2_ncu_full_tables.txt
cp gpurun_out/profiles_r2/r2_conv_tc_res4_stalls.txt profiles/ 2>/dev/null
python tools/ncu_kernel_summary.py gpurun_out/r2_ncu_launches_split_b8.csv > profiles/r2_ncu_launches_split_b8_summary.txt
cp gpurun_out/r2_ncu_launches_split_b8.csv profiles/
python tools/ncu_dram_summary.py gpurun_out/r2_ncu_launches_split_b8.csv profiles/r2_conv_tc_dram_split_b8.json split
for n in r2_effdet_d7_launches r2_osnet_launches; do python tools/ncu_kernel_summary.py gpurun_out/$n.csv > profiles/${n}_summary.txt; done
for f in r2_bench_split.json r2_bench_fp16.json r2_bench_reference_arm.json r2_layers_split_b8.json r2_layers_fp16_b8.json r2_cudnn_layers_b8.jsonl r2_widen_timing.jsonl r2_tracker_bench.jsonl r2_aux_engines.jsonl; do
  [ -f gpurun_out/$f ] && cp gpurun_out/$f profiles/$f
done
cp gpurun_out/baseline_parity.jsonl profiles/r2_baseline_parity.jsonl
cp gpurun_out/r2_gpu_tests.log profiles/r2_gpu_tests.log
( cd tools && python cudnn_table.py ../profiles/r2_cudnn_layers_b8.jsonl ../profiles/r2_layers_split_b8.json ../profiles/r2_layers_fp16_b8.json ) > profiles/r2_cudnn_vs_ours.md
{ python tools/layer_report.py profiles/r2_layers_split_b8.json 40; echo; python tools/layer_report.py profiles/r2_layers_fp16_b8.json 35; } > profiles/r2_layer_table.txt
cat profiles/r2_ncu_full_tables.txt | head -90
