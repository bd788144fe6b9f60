"""Summarise a bench.py --profile-json file: per-layer-shape time, algorithmic TFLOP/s and GB/s."""
import collections
import json
import sys


def main(path, top=22):
    d = json.load(open(path))
    steps = d["steps"]
    tot = sum(s["ms"] for s in steps)
    print("precision=%s batch=%d  total %.3f ms (eager, CUDA events per launch group)" % (d["precision"], d["batch"], tot))
    agg = collections.OrderedDict()
    for s in steps:
        k = s["name"].split(" [")[1].rstrip("]") if " [" in s["name"] else s["name"]
        a = agg.setdefault(k, [0.0, 0.0, 0.0, 0])
        a[0] += s["ms"]; a[1] += s["flops"]; a[2] += s["bytes"]; a[3] += 1
    print("%8s %5s %4s %9s %9s  %s" % ("ms", "%", "n", "TFLOP/s", "GB/s", "layer shape (HxWxCin RxS/stride dil ->Cout)"))
    for k, (ms, fl, by, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%8.3f %5.1f %4d %9.1f %9.0f  %s" % (ms, 100 * ms / tot, n, fl / (ms * 1e-3) / 1e12 if ms else 0,
                                                    by / (ms * 1e-3) / 1e9 if ms else 0, k))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 22)
