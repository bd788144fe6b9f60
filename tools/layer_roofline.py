"""Per-layer floors of a bench.py --profile-json table against the bounds of DESIGN.md 4.3a/4.3b.

For every conv launch group: the tile the plan uses (128 x block_n, block_n as conv_tc.cu: pick_block_n), its tile-wave
quantisation on 148 SMs, the tensor-pipe time of its MMA work at the given SM clock, the cap that operand traffic through
shared memory puts on pipe activity for that tile (TMA fill + tensor-core reads against 128 B/clk/SM), the HBM time of its
algorithmic bytes at the measured copy peak -- and how far the measured time is from the larger of the two floors.
Usage: python tools/layer_roofline.py profiles/r2_layers_split_b8.json [sm_mhz=1680] [top=24]"""
import collections
import json
import math
import re
import sys

SMS = 148
FLOP_PER_CLK_SM = 8192.0      # dense fp16, fp32 accumulate
SMEM_B_PER_CLK = 128.0
HBM_GBS = 6576.0              # MEASURED_PEAKS.json copy bandwidth


def pick_block_n(cout_pad, split):
    cap = 128 if split else 256
    if cout_pad <= cap:
        return cout_pad
    for bn in range(cap, 15, -16):
        if cout_pad % bn == 0:
            return bn
    return 16


def tile_cap(block_n, split):
    """Pipe-activity cap from shared-memory traffic for a 128 x block_n tile, and the fill rate it needs at 100 %."""
    planes = 2 if split else 1
    mmas = 3 if split else 1
    written = (128 + block_n) * 128.0 * planes                 # bytes per 64-deep K-block
    read = 4 * mmas * (128 * 32.0 + block_n * 32.0)            # four K-steps, A tile + B tile per MMA
    clk = 4 * mmas * (block_n / 2.0)                           # a 128 x N x 16 MMA takes N/2 cycles
    return min(1.0, SMEM_B_PER_CLK / ((written + read) / clk)), written / clk


def main(path, sm_mhz=1680.0, top=24):
    d = json.load(open(path))
    split = d["precision"] == "split"
    mmas = 3 if split else 1
    agg = collections.OrderedDict()
    for s in d["steps"]:
        m = re.search(r"\[(\d+)x(\d+)x(\d+) (\d+)x(\d+)/(\d+) d(\d+) ->(\d+)\]", s["name"])
        key = s["name"].split(" [")[1].rstrip("]") if " [" in s["name"] else s["name"]
        a = agg.setdefault(key, dict(ms=0.0, n=0, mma_ms=0.0, cap_ms=0.0, capq_ms=0.0, hbm_ms=0.0, bn=0, cap=0.0, waves=0.0))
        a["ms"] += s["ms"]
        a["n"] += 1
        a["hbm_ms"] += s["bytes"] / (HBM_GBS * 1e9) * 1e3
        if not m or s["flops"] <= 0:
            continue
        cin, r, sw, cout = int(m.group(3)), int(m.group(4)), int(m.group(5)), int(m.group(8))
        k = r * sw * cin
        rows = s["flops"] / (2.0 * k * cout)
        cout_pad = (cout + 15) // 16 * 16
        bn = pick_block_n(cout_pad, split)
        tiles = math.ceil(rows / 128.0) * (cout_pad // bn)
        waves = tiles / float(SMS)
        quant = math.ceil(waves) / waves
        cap, _ = tile_cap(bn, split)
        # MMA work of the padded tiles the kernel really issues
        mma_flop = 2.0 * math.ceil(rows / 128.0) * 128 * k * cout_pad * mmas
        mma_ms = mma_flop / (SMS * FLOP_PER_CLK_SM * sm_mhz * 1e6) * 1e3
        a["mma_ms"] += mma_ms
        a["cap_ms"] += mma_ms / cap
        a["capq_ms"] += mma_ms / cap * quant
        a["bn"], a["cap"], a["waves"] = bn, cap, waves
    tot = sum(a["ms"] for a in agg.values())
    print("precision=%s batch=%d  total %.3f ms; SM clock %.0f MHz; shared memory %d B/clk/SM; HBM %.0f GB/s" %
          (d["precision"], d["batch"], tot, sm_mhz, SMEM_B_PER_CLK, HBM_GBS))
    print("%8s %4s %5s %6s %5s | %8s %8s %8s %8s | %6s %6s  %s" %
          ("ms", "n", "bn", "waves", "cap", "mma100", "mma@cap", "x quant", "hbm", "f_cap", "f_capq", "layer"))
    fl_tot = 0.0
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])[:top]:
        floor = max(a["cap_ms"], a["hbm_ms"])
        floorq = max(a["capq_ms"], a["hbm_ms"])
        print("%8.3f %4d %5s %6s %5s | %8.3f %8.3f %8.3f %8.3f | %6.2f %6.2f  %s" %
              (a["ms"], a["n"], a["bn"] or "-", ("%.2f" % a["waves"]) if a["waves"] else "-",
               ("%.2f" % a["cap"]) if a["cap"] else "-", a["mma_ms"], a["cap_ms"], a["capq_ms"], a["hbm_ms"],
               floor / a["ms"] if a["ms"] else 0, floorq / a["ms"] if a["ms"] else 0, key))
    for a in agg.values():
        fl_tot += max(a["capq_ms"], a["hbm_ms"])
    capsum = sum(max(a["cap_ms"], a["hbm_ms"]) for a in agg.values())
    print("sum of floors: %.3f ms without / %.3f ms with tile-wave quantisation = %.0f %% / %.0f %% of the measured %.3f ms" %
          (capsum, fl_tot, 100 * capsum / tot, 100 * fl_tot / tot, tot))
    print("columns: mma100 = tensor-pipe time of the issued MMAs at 100 % activity; mma@cap = that / cap (cap = 128 B/clk over the tile's"
          " fill + operand-read bytes per MMA cycle); x quant = with ceil(waves)/waves; hbm = algorithmic bytes at the copy peak;"
          " f_cap, f_capq = max(floor) / measured")


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 1680.0, int(sys.argv[3]) if len(sys.argv) > 3 else 24)
