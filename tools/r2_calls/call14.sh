#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for kb in 0 2 4; do
  echo "=== B2_EFFDET_ACC_KB=$kb"
  rm -f gpurun_out/baseline_parity.jsonl
  B2_EFFDET_ACC_KB=$kb timeout 400 python -m pytest tests/test_effdet_gpu.py -q -x --timeout=300 2>&1 | tail -3
  python - <<'PY'
import json
for l in open("gpurun_out/baseline_parity.jsonl"):
    r = json.loads(l)
    print({k: r[k] for k in ("c_rel_exact", "c_rel_oracle32_exact", "gpu_to_exact", "oracle32_to_exact", "gpu_to_oracle32", "prob_maxabs") if k in r})
PY
  B2_EFFDET_ACC_KB=$kb timeout 200 python tools/gpu_aux_timing.py 2>/dev/null | head -1 | cut -c1-220
done
