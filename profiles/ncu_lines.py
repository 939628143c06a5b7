#!/usr/bin/env python3
"""Per-CUDA-source-line instruction and stall-sample shares from
`ncu -i X.ncu-rep --page source --csv --print-source cuda,sass -k regex:<kernel>`."""
import csv
import sys


def main(fn, top=40):
    rows = list(csv.reader(open(fn)))
    out, hdr, cur = [], None, ""
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            ie, ns = hdr.index("Instructions Executed"), hdr.index("# Samples")
            continue
        if hdr and r[0] not in ("", "Function Name") and len(r) > ie:
            try:
                out.append((cur, int(r[0]), int(r[ie]), int(r[ns]), r[1]))
            except ValueError:
                pass
    tot = sum(o[2] for o in out) or 1
    tots = sum(o[3] for o in out) or 1
    print("total warp instructions %d, samples %d" % (tot, tots))
    for o in sorted(out, key=lambda x: -(x[2] / tot + x[3] / tots))[:top]:
        print("%-14s %4d  inst %5.1f%%  samples %5.1f%%  %s" % (o[0], o[1], 100.0 * o[2] / tot, 100.0 * o[3] / tots, o[4].strip()[:110]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
