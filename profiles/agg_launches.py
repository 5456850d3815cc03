#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import re
import sys


def main(path, top=40):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        try:
            t = float(row["Metric Value"].replace(",", ""))
        except (ValueError, KeyError):
            continue
        unit = row["Metric Unit"]
        t = t / 1e3 if unit == "ns" else t * 1e3 if unit == "ms" else t
        name = re.sub(r"\(.*", "", row["Kernel Name"])[:64]
        a = agg.setdefault(name, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += t
        a[2] = max(a[2], t)
    tot = sum(a[1] for a in agg.values())
    print("%-66s %5s %12s %10s %6s" % ("kernel", "n", "total_us", "max_us", "share"))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%-66s %5d %12.1f %10.1f %5.1f%%" % (k, a[0], a[1], a[2], 100 * a[1] / tot))
    print("total_us %.1f" % tot)


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
