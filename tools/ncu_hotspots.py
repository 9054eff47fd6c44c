#!/usr/bin/env python
"""Top stall locations of an `ncu --page source --csv` dump (SASS view).
usage: ncu -i prof.ncu-rep --page source --csv > src.csv; python tools/ncu_hotspots.py src.csv [N]"""
import csv
import sys


def main(path, top=25):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    ia, isrc, isamp, iexec = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    data = []
    for n, r in enumerate(rows[2:]):
        try:
            data.append((int(r[isamp]), n, r))
        except (ValueError, IndexError):
            pass
    total = sum(d[0] for d in data)
    print(f"total samples {total}, instructions {len(data)}")
    for s, n, r in sorted(data, reverse=True)[:top]:
        st = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:3]
        print(f"{100*s/total:5.1f}%  #{n:5d} exec={r[iexec]:>9}  {r[isrc].strip()[:70]:70s}  {st}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
