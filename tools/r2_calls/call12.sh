#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/baseline_parity.jsonl
timeout 1200 python -m pytest tests -m gpu -x -q --timeout=600 --durations=3 2>&1 | tee gpurun_out/r2_gpu_tests.log | tail -8
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
NCU="ncu --profile-from-start off --clock-control none"
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.max.pct_of_peak_sustained_active,sm__cycles_elapsed.avg,dram__bytes_read.sum,dram__bytes_write.sum
for v in dyn static; do
  [ $v = static ] && export B2_STATIC_SCHED=1
  timeout 300 $NCU --metrics $M -k regex:conv_tc --launch-skip 28 -c 3 --csv --log-file gpurun_out/r2_res4_${v}.csv python tools/ncu_pass.py split 8 > /dev/null 2>&1
  echo "== $v"; python - $v <<'PY'
import csv, sys
rows = list(csv.DictReader(l for l in open("gpurun_out/r2_res4_%s.csv" % sys.argv[1]) if not l.startswith("==")))
per = {}
for r in rows:
    per.setdefault(r["ID"], {})[r["Metric Name"]] = r["Metric Value"]
for k, m in per.items():
    print(k, {a.split(".")[0][-28:] + "." + a.split(".")[1]: b for a, b in m.items()})
PY
done
unset B2_STATIC_SCHED
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_split.json 2> gpurun_out/r2_bench_split.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_bench_split.json"))
for k in ("value","ms_per_step","e2e","sustained","stream_c1","roofline","clocks"): print(k, d.get(k))
PY
B2_STATIC_SCHED=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-stream --sustained-seconds 0 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('static: value', d['value'], 'frac', d['roofline']['frac'])"
