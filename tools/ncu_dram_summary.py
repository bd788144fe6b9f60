"""ncu CSV (dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum per conv_tc_kernel launch)
-> profiles/r1_conv_tc_dram_<precision>_b8.json, read by bench.py for roofline.traffic."""
import csv
import json
import sys


def main(path, out, precision):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    per = {}
    for r in rows:
        if "conv_tc_kernel" not in r["Kernel Name"]:
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"].lower()
        if r["Metric Name"].startswith("dram__bytes"):
            scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1)
            per.setdefault(r["ID"], {}).setdefault("bytes", 0.0)
            per[r["ID"]]["bytes"] += v * scale
        elif r["Metric Name"].startswith("gpu__time_duration"):
            scale = {"ns": 1e-9, "nsecond": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3}.get(unit, 1e-9)
            per.setdefault(r["ID"], {})["seconds"] = v * scale
    n = len(per)
    total_b = sum(p.get("bytes", 0) for p in per.values())
    total_s = sum(p.get("seconds", 0) for p in per.values())
    json.dump({"precision": precision, "batch": 8, "launches": n, "dram_bytes_total": total_b,
               "seconds_total_serialised": total_s, "note": "ncu --clock-control none, cold-cache serialised replay"},
              open(out, "w"), indent=1)
    print("launches", n, "dram GB", total_b / 1e9, "ms", total_s * 1e3)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3])
