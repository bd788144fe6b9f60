"""Compact per-launch table from an `ncu --page raw --csv` export: duration, DRAM bytes and GB/s, L2 hit rate,
tensor-pipe activity, issue activity.  python tools/ncu_raw_table.py profiles/x_raw.csv"""
import csv
import re
import sys

COLS = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "rdMB"), ("dram__bytes_write.sum", "wrMB"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("lts__t_sector_hit_rate.pct", "L2hit%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%act"),
        ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("launch__grid_size", "grid"), ("launch__block_size", "block")]


def num(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return float("nan")


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    kn = idx["Kernel Name"]
    print("%-34s" % "kernel" + "".join("%12s" % n for _, n in COLS) + "   GB/s")
    for r in data:
        name = re.sub(r"\(.*", "", r[kn]).replace("void ", "").replace("b2::<unnamed>::", "")[:33]
        vals = []
        for m, _ in COLS:
            if m not in idx:
                vals.append(float("nan"))
                continue
            v = num(r[idx[m]])
            u = units[idx[m]].lower()
            if m.startswith("dram__bytes"):
                v *= {"byte": 1e-6, "kbyte": 1e-3, "mbyte": 1.0, "gbyte": 1e3}.get(u, 1e-6)
            if m.startswith("gpu__time"):
                v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(u, 1.0)
            vals.append(v)
        gbs = (vals[1] + vals[2]) * 1e6 / (vals[0] * 1e-6) / 1e9 if vals[0] == vals[0] and vals[0] > 0 else float("nan")
        print("%-34s" % name + "".join("%12.2f" % v for v in vals) + "%9.0f" % gbs)


if __name__ == "__main__":
    main(sys.argv[1])
