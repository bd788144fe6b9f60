#!/bin/bash
# final verification of the tree with the output TMA warps: whole -m gpu suite, smoke, bench line + per-layer table,
# res4 ncu capture (slimmed on the box), launch list, fp16 bench
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/profiles_r2
rm -f gpurun_out/baseline_parity.jsonl
timeout 560 python -m pytest tests -m gpu -x -q --timeout=300 --durations=4 2>&1 | tee gpurun_out/r2_gpu_tests.log | tail -10
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 200 python bench.py --steps 20 --warmup 3 --profile-json gpurun_out/r2_layers_split_b8.json > gpurun_out/r2_bench_split.json 2> gpurun_out/r2_bench_split.err; echo "bench split rc=$?"; tail -c 600 gpurun_out/r2_bench_split.json
NCU="ncu --profile-from-start off --clock-control none"
timeout 150 $NCU --set full --import-source on -k regex:conv_tc --launch-skip 28 -c 3 -f -o gpurun_out/r2_conv_tc_res4_block1_split python tools/ncu_pass.py split 8 > gpurun_out/ncu_conv.log 2>&1; echo "ncu conv rc=$?"
n=r2_conv_tc_res4_block1_split
if [ -f gpurun_out/$n.ncu-rep ]; then
  ncu -i gpurun_out/$n.ncu-rep --page raw --csv > /tmp/${n}_full.csv 2>/dev/null
  python tools/ncu_raw_slim.py /tmp/${n}_full.csv gpurun_out/profiles_r2/${n}_raw.csv
  ncu -i gpurun_out/$n.ncu-rep --page source --csv --print-source sass > /tmp/conv_src.csv 2>/dev/null
  python tools/ncu_stall_roles.py /tmp/conv_src.csv > gpurun_out/profiles_r2/r2_conv_tc_res4_stalls.txt 2>&1
  rm -f gpurun_out/$n.ncu-rep
fi
timeout 150 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file gpurun_out/r2_ncu_launches_split_b8.csv python tools/ncu_pass.py split 8 > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 100 python bench.py --steps 20 --warmup 3 --precision fp16 --no-cpu-baseline --no-stream --sustained-seconds 0 --profile-json gpurun_out/r2_layers_fp16_b8.json > gpurun_out/r2_bench_fp16.json 2> gpurun_out/r2_bench_fp16.err; echo "bench fp16 rc=$?"; tail -c 300 gpurun_out/r2_bench_fp16.json
