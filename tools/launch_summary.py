"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel: python tools/launch_summary.py file.csv"""
import collections
import csv
import sys


def main(path, top=25):
    with open(path) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    n = 0
    for row in r:
        name = row[ki].split("(")[0]
        v = float(row[vi].replace(",", ""))
        a = agg[name]
        a[0] += 1
        a[1] += v
        a[2] = max(a[2], v)
        n += 1
    tot = sum(a[1] for a in agg.values())
    print(f"launches {n}  total {tot / 1e6:.3f} ms (cold-cache, serialised)")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{k[:62]:62s} n={a[0]:6d} sum={a[1] / 1e6:9.3f} ms  max={a[2] / 1e3:9.1f} us  share={a[1] / tot:.3f}")


if __name__ == "__main__":
    main(sys.argv[1])
