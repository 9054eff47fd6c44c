#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.

usage: python tools/summarise_launches.py gpurun_out/launches.csv > profiles/rNN_launches.md
The per-launch times are cold-cache and serialised (ncu replays each kernel alone):
compare SHARES with bench.py's CUDA-event stage times, not absolutes.
"""
import collections
import csv
import sys


def main(path):
    lines = [l for l in open(path) if l.startswith('"')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0]
        a = agg.setdefault(name, [0, 0.0, row["Grid Size"], row["Block Size"]])
        a[0] += 1
        a[1] += float(row["Metric Value"])
    tot = sum(a[1] for a in agg.values())
    print("| kernel | launches | total ms | share | grid | block |")
    print("|---|---:|---:|---:|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k[:90]}` | {a[0]} | {a[1] / 1e6:.3f} | {100 * a[1] / tot:.1f}% | {a[2]} | {a[3]} |")
    print(f"\ntotal {tot / 1e6:.3f} ms over {sum(a[0] for a in agg.values())} launches")


if __name__ == "__main__":
    main(sys.argv[1])
