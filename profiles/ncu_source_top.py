#!/usr/bin/env python3
"""Summarise an `ncu -i X.ncu-rep --page source --csv` dump: stall mix and the hottest source/SASS lines."""
import csv
import sys


def main(fn, top=24):
    rows = list(csv.reader(open(fn)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    print("#", ",".join(rows[0][:2]) if h else "")
    hdr = rows[h]
    si, ns, ie = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stalls = [i for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
    tot, agg, data, inst = 0, {}, [], 0
    for r in rows[h + 1:]:
        try:
            s = int(r[ns])
        except (ValueError, IndexError):
            continue
        tot += s
        try:
            inst += int(r[ie])
        except ValueError:
            pass
        for i in stalls:
            try:
                agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i])
            except ValueError:
                pass
        data.append((s, r[si], r[ie]))
    print("total samples %d, warp instructions executed %d" % (tot, inst))
    for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]:
        print("   %-25s %6.1f%%" % (k, 100.0 * v / max(1, tot)))
    for s, src, n in sorted(data, key=lambda x: -x[0])[:top]:
        print("   %5.1f%% %10s  %s" % (100.0 * s / max(1, tot), n, src[:120]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 24)
