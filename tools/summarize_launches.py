#!/usr/bin/env python
"""ncu `--metrics gpu__time_duration.sum --csv` launch list -> per-kernel summary (profiles/*.txt)."""
import collections
import csv
import sys


def main(path, skip_pack=True):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0]
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print("# %s: %d launches, %.1f us total (cold-cache, serialised: compare SHARES)" % (path, sum(v[0] for v in agg.values()), tot))
    print("%-58s %6s %12s %10s %7s" % ("kernel", "n", "total_us", "avg_us", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-58s %6d %12.1f %10.2f %6.1f%%" % (k[:58], v[0], v[1], v[1] / v[0], 100 * v[1] / tot))


if __name__ == "__main__":
    main(sys.argv[1])
