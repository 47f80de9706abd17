#!/usr/bin/env python
"""Per-SASS-instruction stall samples from an `ncu --page source --csv` dump: hot instructions and region totals.

    ncu -i rep.ncu-rep --page source --csv > src.csv ; python tools/ncu_source_hot.py src.csv [min_share]
"""
import csv
import sys


def main(path, min_share=0.002):
    rows = list(csv.reader(open(path)))
    hdr, data = rows[1], rows[2:]
    ia, isrc, isamp, iex = (hdr.index(k) for k in ("Address", "Source", "# Samples", "Instructions Executed"))
    stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[isamp]) for r in data)
    base = int(data[0][ia], 16)
    mx = max(int(r[iex]) for r in data)
    print("total samples", tot, "max executed", mx)
    agg = {}
    for r in data:
        a = int(r[ia], 16) - base
        s, ex = int(r[isamp]), int(r[iex])
        for i in stall:
            agg[hdr[i][6:]] = agg.get(hdr[i][6:], 0) + int(r[i])
        if s > tot * min_share or ex > mx * 0.5:
            st = {hdr[i][6:]: int(r[i]) for i in stall if int(r[i]) > 0}
            top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
            print(f"{a:6x} {ex:12d} {s:7d} {100 * s / tot:5.2f}% {r[isrc].strip()[:64]:64s} {top}")
    print({k: round(100 * v / tot, 2) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 0.002)
