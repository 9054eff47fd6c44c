#!/usr/bin/env python
"""Per-kernel summary of an ncu report (raw page): duration, DRAM traffic and throughput, tensor-pipe and
issue utilisation, registers.  Launches of the same kernel are averaged.

usage: ncu -i prof.ncu-rep --page raw --csv | python tools/ncu_summary.py [--json out.json --frames N]
"""
import collections
import csv
import json
import sys

COLS = {
    "dur_us": "gpu__time_duration.sum",
    "dram_rd": "dram__bytes_read.sum",
    "dram_wr": "dram__bytes_write.sum",
    "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "tensor_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "issue_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "regs": "launch__registers_per_thread",
    "warps_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
}
SCALE = {"nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6,
         "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    rows = list(csv.reader(sys.stdin))
    hdr, units = rows[0], rows[1]
    idx = {k: hdr.index(v) for k, v in COLS.items() if v in hdr}
    iname = hdr.index("Kernel Name")
    agg = collections.OrderedDict()
    for r in rows[2:]:
        name = r[iname].split("(")[0].replace("void ", "").replace("pmb::", "")
        a = agg.setdefault(name, collections.defaultdict(float))
        a["n"] += 1
        for k, i in idx.items():
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                continue
            a[k] += v * SCALE.get(units[i], 1.0)
    out = {}
    print("| kernel | launches | avg µs | DRAM MB / launch | DRAM % of peak | tensor pipe % | issue % | regs |")
    print("|---|---:|---:|---:|---:|---:|---:|---:|")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["dur_us"]):
        n = a["n"]
        mb = (a["dram_rd"] + a["dram_wr"]) / n / 1e6
        print(f"| `{name[:60]}` | {int(n)} | {a['dur_us'] / n:.1f} | {mb:.1f} | {a['dram_pct'] / n:.1f} | "
              f"{a['tensor_pct'] / n:.1f} | {a['issue_pct'] / n:.1f} | {a['regs'] / n:.0f} |")
        out[name] = {"launches": int(n), "avg_us": a["dur_us"] / n, "dram_bytes_per_launch": (a["dram_rd"] + a["dram_wr"]) / n,
                     "dram_pct": a["dram_pct"] / n, "tensor_pct": a["tensor_pct"] / n, "issue_pct": a["issue_pct"] / n}
    if "--json" in sys.argv:
        path = sys.argv[sys.argv.index("--json") + 1]
        frames = int(sys.argv[sys.argv.index("--frames") + 1]) if "--frames" in sys.argv else None
        json.dump({"frames": frames, "kernels": out}, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
