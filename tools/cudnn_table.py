"""Joins the cuDNN library baseline (tools/cudnn_layer_baseline.py -> jsonl) with this repo's per-layer tables
(bench.py --profile-json, split and fp16 precision) into one markdown table: ms per launch and algorithmic TF/s.
  python tools/cudnn_table.py cudnn.jsonl layers_split.json [layers_fp16.json] > profiles/r2_cudnn_vs_ours.md"""
import collections
import json
import re
import sys

from cudnn_layer_baseline import LAYERS


def ours(path):
    d = json.load(open(path))
    out = collections.OrderedDict()
    for s in d["steps"]:
        m = re.search(r"\[(\d+)x(\d+)x(\d+) (\d+)x(\d+)/(\d+) d(\d+) ->(\d+)\]", s["name"])
        if not m:
            continue
        k = tuple(int(v) for v in m.groups())
        a = out.setdefault(k, [0.0, 0.0, 0])
        a[0] += s["ms"]; a[1] += s["flops"]; a[2] += 1
    return out


def main():
    cud = collections.defaultdict(dict)
    for line in open(sys.argv[1]):
        r = json.loads(line)
        cud[r["layer"]][r["mode"]] = r
    tabs = [ours(p) for p in sys.argv[2:]]
    names = ["split (fp32-class)", "fp16"][:len(tabs)]
    print("| layer (batch 8) | launches/pass | " + " | ".join("ours %s: ms/launch (TF/s)" % n for n in names) +
          " | cuDNN fp16 | cuDNN TF32 | cuDNN fp32 |")
    print("|---|---|" + "---|" * (len(tabs) + 3))
    for name, H, W, ci, co, k, s, dl, cnt in LAYERS:
        cells = []
        n_l = 0
        for t in tabs:
            hit = [v for kk, v in t.items() if kk[2] == ci and kk[3] == k and kk[5] == s and kk[6] == dl and kk[7] == co
                   and abs(kk[0] - H) <= 4 and abs(kk[1] - W) <= 4]
            if hit:
                keys = [kk for kk in t if kk[2] == ci and kk[3] == k and kk[5] == s and kk[6] == dl and kk[7] == co
                        and abs(kk[0] - H) <= 4 and abs(kk[1] - W) <= 4]
                best = min(range(len(hit)), key=lambda i: abs(keys[i][0] - H) + abs(keys[i][1] - W))
                ms, fl, n = hit[best]
                n_l = n
                cells.append("%.4f (%.0f)" % (ms / n, fl / (ms * 1e-3) / 1e12))
            else:
                cells.append("-")
        c = cud.get(name, {})
        cc = ["%.4f (%.0f)" % (c[m]["ms"], c[m]["tflops"]) if m in c else "-" for m in ("fp16", "tf32", "fp32")]
        print("| %s | %s | %s | %s |" % (name, n_l or "-", " | ".join(cells), " | ".join(cc)))


if __name__ == "__main__":
    main()
