#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/baseline_parity.jsonl
echo "=== dynamic tile scheduler: conv + pipeline + reid + effdet quick tests"
timeout 300 python -m pytest tests/test_conv_gpu.py tests/test_pipeline_gpu.py -x -q --timeout=120 2>&1 | tail -6
[ ${PIPESTATUS[0]} -eq 0 ] || { echo "STOP: quick tests failed / hung"; B2_STATIC_SCHED=1 timeout 300 python -m pytest tests/test_conv_gpu.py tests/test_pipeline_gpu.py -x -q --timeout=120 2>&1 | tail -4; exit 1; }
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 --durations=4 2>&1 | tail -15
echo "=== margins"; python - <<'PY'
import json
for l in open("gpurun_out/baseline_parity.jsonl"):
    r = json.loads(l)
    keys = ("config", "seed", "image", "final_set_dist", "gpu_to_exact", "oracle32_to_exact", "gpu_to_oracle32", "prob_maxabs", "c_rel_exact", "c_rel_oracle32_exact")
    print({k: (round(v, 7) if isinstance(v, float) else v) for k, v in r.items() if k in keys})
PY
tools/ab_run.sh dyn: static:B2_STATIC_SCHED=1 dynshortk2:B2_ACC_KB_SHORTK=2 pdl-dyn: pdl-dynshortk2:B2_ACC_KB_SHORTK=2 dyn2: static2:B2_STATIC_SCHED=1
echo "=== accuracy of shortk2 + delta2"
B2_ACC_KB_SHORTK=2 timeout 300 python tools/gpu_pipeline_probe.py 720 1280 tcgen05 split 3,4,23,3 2 2>&1 | grep EXACT | grep -v "c[23] rel"
