#!/bin/bash
# final verification of the tree that carries the opt-in CTA-pair path (default path = the verified single-CTA kernel,
# recompiled): whole -m gpu suite + smoke; then, budget permitting, pairs with two K-blocks per chunk against the default
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/baseline_parity.jsonl
timeout -k 5 330 python -m pytest tests -m gpu -x -q --timeout=200 --durations=3 2>&1 | tee gpurun_out/r2_gpu_tests.log | tail -8
timeout -k 5 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for v in pairkb2 base; do
  unset B2_PAIR B2_PAIR_ACC_KB
  if [ "$v" = "pairkb2" ]; then export B2_PAIR=8 B2_PAIR_ACC_KB=2; fi
  timeout -k 5 80 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-stream --sustained-seconds 0 \
     --profile-json gpurun_out/layers_$v.json > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
  echo "== $v rc=$?"; python -c "
import json
try:
    d=json.loads(open('gpurun_out/bench_$v.json').read().strip().splitlines()[-1]); print('   %.1f FPS  %.3f ms/step  e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']))
except Exception as e: print('   no bench line', e)
"
done
