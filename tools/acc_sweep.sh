mkdir -p gpurun_out
for kb in 1 2 4; do
  echo "=== ACC_KB=$kb"
  B2_ACC_KB=$kb timeout 600 python tools/gpu_pipeline_probe.py 720 1280 tcgen05 split > gpurun_out/pipe_acc_kb$kb.log 2>&1
  grep -E "c[45] rel|proposals gpu|final gpu" gpurun_out/pipe_acc_kb$kb.log
  B2_ACC_KB=$kb timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_acc_kb$kb.log 2>&1
  tail -1 gpurun_out/bench_acc_kb$kb.log | cut -c1-140
done
