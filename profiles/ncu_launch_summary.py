#!/usr/bin/env python3
"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total time, share."""
import collections
import csv
import sys


def main(fn):
    rows = list(csv.reader(l for l in open(fn) if l.startswith('"')))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        if r[ui] == "ns":
            v /= 1000.0
        elif r[ui] == "ms":
            v *= 1000.0
        a = agg.setdefault(r[ki].split("(")[0], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("%-72s %8s %12s %7s" % ("kernel", "launches", "total_us", "share"))
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%-72s %8d %12.1f %6.1f%%" % (k[:72], a[0], a[1], 100 * a[1] / tot))
    print("%-72s %8d %12.1f" % ("TOTAL", sum(a[0] for a in agg.values()), tot))


if __name__ == "__main__":
    main(sys.argv[1])
