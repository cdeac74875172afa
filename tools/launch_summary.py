"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel. Usage: launch_summary.py file.csv"""
import collections
import csv
import re
import sys


def load(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = []
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v * 1e6 if u == "s" else v
        rows.append((re.sub(r"\(.*", "", r["Kernel Name"]), r.get("Grid Size", ""), v))
    return rows


if __name__ == "__main__":
    rows = load(sys.argv[1])
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for k, g, v in rows:
        agg[k][0] += 1
        agg[k][1] += v
        agg[k][2] = max(agg[k][2], v)
    tot = sum(a[1] for a in agg.values())
    print(f"{len(rows)} launches, {tot / 1e3:.2f} ms total (cold-cache, serialised: compare shares)")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{k[:70]:70s} n={a[0]:5d} total={a[1] / 1e3:9.3f} ms mean={a[1] / a[0]:9.1f} us max={a[2]:9.1f} us {100 * a[1] / tot:5.1f}%")
