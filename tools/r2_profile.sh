#!/bin/bash
# Round-2 measurement pass for ONE gpurun call (one GPU): ncu launch list of the bench workload, full captures of the res4
# convs, of the non-conv kernels of the pass, of one BiFPN cell, one OSNet OSBlock and the small late-round-1 kernels.
# Everything lands in gpurun_out/; tools/r2_profile_collect.sh (here, afterwards) turns the .ncu-rep files into the CSV /
# text summaries under profiles/.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
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
ls -la gpurun_out/*.ncu-rep
