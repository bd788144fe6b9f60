#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/baseline_parity.jsonl
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 --durations=4 2>&1 | tail -25
echo "=== margins"; python - <<'PY'
import json
for l in open("gpurun_out/baseline_parity.jsonl"):
    r = json.loads(l)
    keys = ("config", "seed", "image", "final", "final_ref", "final_set_dist", "gpu_to_exact", "oracle32_to_exact", "gpu_to_oracle32", "prob_maxabs", "lvl_set_dist", "proposal_set_dist", "c_rel", "c_rel_exact", "c_rel_oracle32_exact", "box_rel")
    print({k: (round(v, 7) if isinstance(v, float) else v) for k, v in r.items() if k in keys})
PY
echo "=== delta fine sweep"
for d in 7e-8 1e-7 1.2e-7; do
  echo "--- B2_ACC_DELTA=$d"
  B2_ACC_DELTA=$d timeout 300 python tools/gpu_pipeline_probe.py 720 1280 tcgen05 split 3,4,23,3 2 2>&1 | grep EXACT | grep -v "c[234] rel"
done | tee gpurun_out/acc_delta_sweep_fine.txt
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_new.json 2> gpurun_out/bench_new.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_new.json"))
for k in ("value","ms_per_step","e2e","sustained","stream_c1","roofline"): print(k, d.get(k))
PY
