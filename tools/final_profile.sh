#!/bin/bash
# Round-end measurement pass for one gpurun call (one GPU): parity tests, bench lines, ncu launch list and full captures.
# Everything lands in gpurun_out/; copy what is to be judged into profiles/.
#   tools/final_profile.sh [post] [cosine] [variant specs for tools/ab_run.sh ...]
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_conv_gpu.py -x -q 2>&1 | tail -4
[ ${PIPESTATUS[0]} -eq 0 ] || { echo "conv tests failed: stop"; exit 1; }
timeout 420 python -m pytest tests -m gpu -x -q --timeout=150 --durations=5 2>&1 | tail -14
timeout 300 python bench.py --steps 20 --warmup 3 --profile-json gpurun_out/r1_layers_split_b8.json > gpurun_out/r1_bench_split.json 2> gpurun_out/r1_bench_split.err; echo "bench split rc=$?"
NCU="ncu --profile-from-start off --clock-control none"
# res4 block1 conv1 / conv2 / conv3: the 29th..31st conv_tc launches of an eager pass
timeout 300 $NCU --set full --import-source on -k regex:conv_tc --launch-skip 28 -c 3 -f -o gpurun_out/r1_conv_tc_res4_block1_split python tools/ncu_pass.py split 8 > gpurun_out/ncu_conv.log 2>&1; echo "ncu conv rc=$?"
timeout 300 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r1_ncu_launches_split_b8.csv python tools/ncu_pass.py split 8 > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 200 python bench.py --steps 20 --warmup 3 --precision fp16 --no-cpu-baseline --profile-json gpurun_out/r1_layers_fp16_b8.json > gpurun_out/r1_bench_fp16.json 2> gpurun_out/r1_bench_fp16.err; echo "bench fp16 rc=$?"
for arg in "$@"; do
  case "$arg" in
    post) timeout 300 $NCU --set full --import-source on -k regex:'rpn_|roialign|head_decode|class_nms|final_topk|stem_pack|maxpool' -c 12 -f -o gpurun_out/r1_post_kernels_split python tools/ncu_pass.py split 8 > gpurun_out/ncu_post.log 2>&1; echo "ncu post rc=$?" ;;
    cosine) timeout 200 ncu --clock-control none --set full --import-source on -k regex:'cosine|conv_tc' -c 6 -f -o gpurun_out/r1_cosine_cost python tools/ncu_cosine.py > gpurun_out/ncu_cosine.log 2>&1; echo "ncu cosine rc=$?" ;;
    *) tools/ab_run.sh "$arg" ;;
  esac
done
ls -la gpurun_out | tail -12
