#!/usr/bin/env python
"""One optimizer step of an ncu launch list (between two k_clamp_adam launches): per-kernel totals + the big launches."""
import collections
import csv
import sys


def main(path, thresh=150.0):
    lines = [l for l in open(path) if not l.startswith("==")]
    seq = []
    for r in csv.DictReader(lines):
        n = r["Kernel Name"].split("(")[0]
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
        seq.append((n, v, r["Grid Size"]))
    idx = [i for i, s in enumerate(seq) if "clamp_adam" in s[0]]
    a, b = idx[0], idx[1]
    agg = collections.OrderedDict()
    tot = 0.0
    for i in range(a + 1, b + 1):
        n, v, g = seq[i]
        tot += v
        k = n[:58]
        agg.setdefault(k, [0, 0.0])
        agg[k][0] += 1
        agg[k][1] += v
        if v > thresh:
            print("%4d %-60s %9.1f %s" % (i - a, n[:60], v, g))
    print("# one step: %d launches, %.1f us (cold-cache, serialised)" % (b - a, tot))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:26]:
        print("%-58s %5d %9.1f %5.1f%%" % (k, v[0], v[1], 100 * v[1] / tot))


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 150.0)
