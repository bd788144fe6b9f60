"""Per-kernel, per-warp-role stall summary from an `ncu --page source --csv --print-source sass` export of conv_tc_kernel
captures: samples of the producer / MMA / epilogue code ranges (split at the role branches by sampled-instruction index) and
their top stall reasons, plus the hottest instructions.  python tools/ncu_stall_roles.py src.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
kernels, cur, hdr = [], None, None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        kernels.append(cur)
        hdr = None
        continue
    if r and r[0] == "Address":
        hdr = r
        cur["hdr"] = r
        continue
    if cur is not None and hdr is not None and r:
        cur["rows"].append(r)
for k, K in enumerate(kernels):
    h = K["hdr"]
    iS, isrc = h.index("# Samples"), h.index("Source")
    stall = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    tot = sum(int(r[iS] or 0) for r in K["rows"])
    print("== kernel %d: %d SASS instructions, %d samples" % (k, len(K["rows"]), tot))
    # role boundaries: the three long waits on mbarriers (BRA after SYNCS try_wait) that lead the sample counts mark the
    # producer / MMA / epilogue loops; report fixed fractions of the instruction index range as a coarse split instead
    n = len(K["rows"])
    for name, a, b in (("first 6% (producer)", 0, int(n * 0.06)), ("6-11% (MMA issuer)", int(n * 0.06), int(n * 0.11)),
                       ("11-98% (epilogue)", int(n * 0.11), int(n * 0.98)), ("tail (exit barrier)", int(n * 0.98), n)):
        agg, cnt = {}, 0
        for r in K["rows"][a:b]:
            cnt += int(r[iS] or 0)
            for i, c in stall:
                agg[c[6:]] = agg.get(c[6:], 0) + int(r[i] or 0)
        top = sorted(((v, c) for c, v in agg.items() if v), reverse=True)[:6]
        print("   %-22s %6d samples  %s" % (name, cnt, ", ".join("%s %d" % (c, v) for v, c in top)))
    hot = sorted(range(n), key=lambda i: -int(K["rows"][i][iS] or 0))[:8]
    for i in sorted(hot):
        r = K["rows"][i]
        print("   [%5d] %5s  %s" % (i, r[iS], r[isrc].strip()[:70]))
