#!/bin/bash
# A/B harness for one gpurun call: pipeline parity tests, then bench lines for the variants named on the command line.
# Usage: tools/ab_run.sh [tests] [name:ENV=V,ENV=V ...]   (a name whose part before "-" matches libb200det_<part>.so runs with that library swapped in)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=object_detection_tracking_b200
for arg in "$@"; do
  if [ "$arg" = "tests" ]; then
    timeout 900 python -m pytest tests/test_pipeline_gpu.py tests/test_backend_gpu.py tests/test_conv_gpu.py -x -q 2>&1 | tail -12
    continue
  fi
  name="${arg%%:*}"; envs="${arg#*:}"; [ "$envs" = "$arg" ] && envs=""
  libn="${name%%-*}"
  swapped=0
  if [ -f $P/libb200det_$libn.so ]; then cp $P/libb200det.so /tmp/new.so; cp $P/libb200det_$libn.so $P/libb200det.so; swapped=1; fi
  ( for kv in ${envs//,/ }; do export "$kv"; done
    timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-json gpurun_out/layers_$name.json \
      > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
    echo "== $name [$envs] rc=$?"
    python - "$name" <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/bench_%s.json" % sys.argv[1]))
    print("   value %.1f FPS  %.3f ms/step  e2e %.1f  launches %s  phases %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("gpu_launches"), {k: round(v, 2) for k, v in d.get("phase_ms", {}).items()}))
except Exception as e:
    print("   no bench line:", e)
PY
    tail -3 gpurun_out/bench_$name.err )
  if [ $swapped = 1 ]; then cp /tmp/new.so $P/libb200det.so; fi
done
