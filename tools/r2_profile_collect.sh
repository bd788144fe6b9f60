#!/bin/bash
# Turns the .ncu-rep files a tools/r2_profile.sh call left in gpurun_out/ into the committed summaries under profiles/.
set -u
cd "$(dirname "$0")/.."
for n in r2_conv_tc_res4_block1_split r2_post_kernels_split r2_bifpn_cell r2_effnet_mbconv r2_osnet_osblock r2_misc_kernels; do
  [ -f gpurun_out/$n.ncu-rep ] || { echo "missing $n"; continue; }
  ncu -i gpurun_out/$n.ncu-rep --page raw --csv > profiles/${n}_raw_full.csv 2>/dev/null
  python tools/ncu_raw_slim.py profiles/${n}_raw_full.csv profiles/${n}_raw.csv && rm profiles/${n}_raw_full.csv
  echo "== $n"; python tools/ncu_raw_table.py profiles/${n}_raw.csv
done > profiles/r2_ncu_full_tables.txt
python tools/ncu_kernel_summary.py gpurun_out/r2_ncu_launches_split_b8.csv > profiles/r2_ncu_launches_split_b8_summary.txt
cp gpurun_out/r2_ncu_launches_split_b8.csv profiles/
python tools/ncu_dram_summary.py gpurun_out/r2_ncu_launches_split_b8.csv profiles/r2_conv_tc_dram_split_b8.json split
for n in r2_effdet_d7_launches r2_osnet_launches; do python tools/ncu_kernel_summary.py gpurun_out/$n.csv > profiles/${n}_summary.txt; done
tail -n +1 profiles/r2_ncu_full_tables.txt | head -80
