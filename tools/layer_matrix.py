"""Side-by-side per-layer-shape times (ms) of several bench.py --profile-json files: python tools/layer_matrix.py a.json b.json ..."""
import collections
import json
import os
import sys


def agg(path):
    d = json.load(open(path))
    out = collections.OrderedDict()
    for s in d["steps"]:
        k = s["name"].split(" [")[1].rstrip("]") if " [" in s["name"] else s["name"]
        out[k] = out.get(k, 0.0) + s["ms"]
    return out


def main(paths, top=18):
    tabs = [agg(p) for p in paths]
    names = [os.path.basename(p).replace("layers_", "").replace(".json", "")[:10] for p in paths]
    keys = sorted(tabs[0], key=lambda k: -tabs[0][k])[:top]
    print(" ".join("%10s" % n for n in names) + "  layer")
    for k in keys:
        print(" ".join("%10.3f" % t.get(k, float("nan")) for t in tabs) + "  " + k)
    print(" ".join("%10.3f" % sum(t.values()) for t in tabs) + "  total (eager)")


if __name__ == "__main__":
    main(sys.argv[1:])
