#!/usr/bin/env python
"""Top source lines by warp-stall samples from `ncu --page source --csv --print-source cuda,sass`."""
import csv
import sys


def main(path, top=25):
    rows = list(csv.reader(open(path)))
    cur_file, out, total = None, [], 0
    hdr = None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < 8 or not r[0].isdigit():
            continue
        if r[2] != "-":          # SASS row, skip (line rows have '-' address)
            continue
        try:
            samples = int(r[4])
        except ValueError:
            continue
        total += samples
        out.append((samples, cur_file, int(r[0]), r[1].strip()[:110], r[7]))
    out.sort(reverse=True)
    print("total samples", total)
    for s, f, ln, src, inst in out[:top]:
        print("%6d %5.1f%%  %s:%d  %s" % (s, 100.0 * s / max(total, 1), f, ln, src))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
