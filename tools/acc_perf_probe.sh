mkdir -p gpurun_out
run() { echo "=== $1"; env $1 timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --profile-json gpurun_out/lp.json > gpurun_out/b.log 2>&1; python - <<'PY'
import json
l=open('gpurun_out/b.log').read().strip().splitlines()[-1]
try:
    d=json.loads(l); print("FPS %.1f  conv frac %.3f" % (d['value'], d['roofline']['frac']))
    s=json.load(open('gpurun_out/lp.json'))['steps']
    for key in ('46x80x256 3x3','184x320x256 3x3','46x80x1024 1x1','46x80x256 1x1'):
        ms=sum(x['ms'] for x in s if key in x['name']); print("   %-22s %.3f ms" % (key, ms))
except Exception as e: print("ERR", e, l[:300])
PY
}
run "B2_NO_ACC=1"
run "B2_ACC_KB=1"
run "B2_ACC_KB=1 B2_ACC_NODRAIN=1"
run "B2_ACC_KB=2"
run "B2_ACC_KB=2 B2_ACC_NODRAIN=1"
