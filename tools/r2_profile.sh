#!/bin/bash
# Round-2 measurement pass for ONE gpurun call (one GPU): ncu launch list of the bench workload, full captures of the res4
# convs, of the non-conv kernels of the pass, of one BiFPN cell, one OSNet OSBlock and the small late-round-1 kernels.
# Everything lands in gpurun_out/; tools/r2_profile_collect.sh (here, afterwards) turns the .ncu-rep files into the CSV /
# text summaries under profiles/.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/baseline_parity.jsonl
timeout 1200 python -m pytest tests -m gpu -x -q --timeout=600 --durations=4 2>&1 | tee gpurun_out/r2_gpu_tests.log | tail -12
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
NCU="ncu --profile-from-start off --clock-control none"
timeout 400 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file gpurun_out/r2_ncu_launches_split_b8.csv python tools/ncu_pass.py split 8 > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 400 $NCU --set full --import-source on -k regex:conv_tc --launch-skip 28 -c 3 -f -o gpurun_out/r2_conv_tc_res4_block1_split python tools/ncu_pass.py split 8 > gpurun_out/ncu_conv.log 2>&1; echo "ncu conv rc=$?"
timeout 400 $NCU --set full --import-source on -k regex:'rpn_|roialign|head_decode|class_nms|final_topk|stem_pack|maxpool' -c 12 -f -o gpurun_out/r2_post_kernels_split python tools/ncu_pass.py split 8 > gpurun_out/ncu_post.log 2>&1; echo "ncu post rc=$?"
# one BiFPN cell of D7 at 1536x1536: the combine / depthwise / pointwise kernels of cell 3 (skip the backbone + 3 cells)
timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2_effdet_d7_launches.csv python tools/effdet_ncu_pass.py efficientdet-d7 1536 1536 split > gpurun_out/ncu_effdet_l.log 2>&1; echo "ncu effdet launches rc=$?"
timeout 600 $NCU --set full --import-source on -k regex:'bifpn_combine|dw3x3_plain' --launch-skip 60 -c 8 -f -o gpurun_out/r2_bifpn_cell python tools/effdet_ncu_pass.py efficientdet-d7 1536 1536 split > gpurun_out/ncu_bifpn.log 2>&1; echo "ncu bifpn rc=$?"
timeout 600 $NCU --set full --import-source on -k regex:'dw_strip|se_partial|se_fc' --launch-skip 20 -c 8 -f -o gpurun_out/r2_effnet_mbconv python tools/effdet_ncu_pass.py efficientdet-d7 1536 1536 split > gpurun_out/ncu_mbconv.log 2>&1; echo "ncu mbconv rc=$?"
# one OSNet OSBlock (batch 64): light-conv depthwise, channel gate, the 1x1 GEMMs
timeout 400 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2_osnet_launches.csv python tools/ncu_reid.py 64 split > gpurun_out/ncu_reid_l.log 2>&1; echo "ncu reid launches rc=$?"
timeout 400 $NCU --set full --import-source on --launch-skip 4 -c 24 -f -o gpurun_out/r2_osnet_osblock python tools/ncu_reid.py 64 split > gpurun_out/ncu_reid.log 2>&1; echo "ncu reid rc=$?"
timeout 400 $NCU --set full --import-source on -k regex:'resize_u8|pair_segmin|mask_select|agg_feat|roialign' -c 10 -f -o gpurun_out/r2_misc_kernels python tools/ncu_misc.py > gpurun_out/ncu_misc.log 2>&1; echo "ncu misc rc=$?"
# convert on the box: gpurun merges at most 64 MiB back, the six .ncu-rep files are 94 MB
mkdir -p gpurun_out/profiles_r2
for n in r2_conv_tc_res4_block1_split r2_post_kernels_split r2_bifpn_cell r2_effnet_mbconv r2_osnet_osblock r2_misc_kernels; do
  [ -f gpurun_out/$n.ncu-rep ] || { echo "missing $n"; continue; }
  ncu -i gpurun_out/$n.ncu-rep --page raw --csv > /tmp/${n}_full.csv 2>/dev/null
  python tools/ncu_raw_slim.py /tmp/${n}_full.csv gpurun_out/profiles_r2/${n}_raw.csv
done
ncu -i gpurun_out/r2_conv_tc_res4_block1_split.ncu-rep --page source --csv --print-source sass > /tmp/conv_src.csv 2>/dev/null
python tools/ncu_stall_roles.py /tmp/conv_src.csv > gpurun_out/profiles_r2/r2_conv_tc_res4_stalls.txt 2>&1
ls -la gpurun_out/*.ncu-rep
rm -f gpurun_out/r2_bifpn_cell.ncu-rep gpurun_out/r2_effnet_mbconv.ncu-rep gpurun_out/r2_osnet_osblock.ncu-rep gpurun_out/r2_misc_kernels.ncu-rep gpurun_out/r2_post_kernels_split.ncu-rep
timeout 300 python bench.py --steps 20 --warmup 3 --profile-json gpurun_out/r2_layers_split_b8.json > gpurun_out/r2_bench_split.json 2> gpurun_out/r2_bench_split.err; echo "bench split rc=$?"
timeout 300 python bench.py --steps 20 --warmup 3 --precision fp16 --no-cpu-baseline --no-stream --sustained-seconds 0 --profile-json gpurun_out/r2_layers_fp16_b8.json > gpurun_out/r2_bench_fp16.json 2> gpurun_out/r2_bench_fp16.err; echo "bench fp16 rc=$?"
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err; echo "reference arm rc=$?"
timeout 200 python tools/cudnn_layer_baseline.py 8 > gpurun_out/r2_cudnn_layers_b8.jsonl 2>/dev/null; echo "cudnn rc=$?"
timeout 200 python tools/gpu_widen_timing.py > gpurun_out/r2_widen_timing.jsonl 2>/dev/null; cat gpurun_out/r2_widen_timing.jsonl
timeout 200 python tools/gpu_tracker_probe.py 2>/dev/null | tail -1 > gpurun_out/r2_tracker_bench.jsonl; cat gpurun_out/r2_tracker_bench.jsonl
timeout 300 python tools/gpu_aux_timing.py > gpurun_out/r2_aux_engines.jsonl 2>/dev/null; cat gpurun_out/r2_aux_engines.jsonl
