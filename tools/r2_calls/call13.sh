#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/baseline_parity.jsonl
timeout 1200 python -m pytest tests -m gpu -x -q --timeout=600 --durations=3 2>&1 | tee gpurun_out/r2_gpu_tests.log | tail -8
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
NCU="ncu --profile-from-start off --clock-control none"
timeout 400 $NCU --set full --import-source on -k regex:conv_tc --launch-skip 28 -c 3 -f -o gpurun_out/r2_conv_tc_res4_block1_split python tools/ncu_pass.py split 8 > gpurun_out/ncu_conv.log 2>&1; echo "ncu conv rc=$?"
mkdir -p gpurun_out/profiles_r2
ncu -i gpurun_out/r2_conv_tc_res4_block1_split.ncu-rep --page raw --csv > /tmp/c_full.csv 2>/dev/null
python tools/ncu_raw_slim.py /tmp/c_full.csv gpurun_out/profiles_r2/r2_conv_tc_res4_block1_split_raw.csv
python tools/ncu_raw_table.py gpurun_out/profiles_r2/r2_conv_tc_res4_block1_split_raw.csv
ncu -i gpurun_out/r2_conv_tc_res4_block1_split.ncu-rep --page source --csv --print-source sass > /tmp/conv_src.csv 2>/dev/null
python tools/ncu_stall_roles.py /tmp/conv_src.csv > gpurun_out/profiles_r2/r2_conv_tc_res4_stalls.txt 2>&1
timeout 400 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file gpurun_out/r2_ncu_launches_split_b8.csv python tools/ncu_pass.py split 8 > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
tools/ab_run.sh base: norange: base2: norange2:
timeout 300 python bench.py --steps 20 --warmup 3 --profile-json gpurun_out/r2_layers_split_b8.json > gpurun_out/r2_bench_split.json 2> gpurun_out/r2_bench_split.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_bench_split.json"))
for k in ("value","ms_per_step","e2e","sustained","stream_c1","roofline","clocks","cpu_baseline"): print(k, d.get(k))
PY
timeout 300 python bench.py --steps 20 --warmup 3 --precision fp16 --no-cpu-baseline --no-stream --sustained-seconds 0 --profile-json gpurun_out/r2_layers_fp16_b8.json > gpurun_out/r2_bench_fp16.json 2> gpurun_out/r2_bench_fp16.err; echo "bench fp16 rc=$?"; python -c "import json; d=json.load(open('gpurun_out/r2_bench_fp16.json')); print('fp16', d['value'], d['roofline']['frac'])"
timeout 300 python tools/gpu_aux_timing.py > gpurun_out/r2_aux_engines.jsonl 2>/dev/null; cat gpurun_out/r2_aux_engines.jsonl
