#!/usr/bin/env python
"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name.
usage: python profiles/aggregate_launches.py gpurun_out/launches.csv [nfirst]"""
import collections
import csv
import sys


def main(path, nfirst=0):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    order = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = row["Kernel Name"][:64]
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
        agg[name][0] += 1
        agg[name][1] += v
        order.append((name, v, row["Grid Size"], row["Block Size"]))
    tot = sum(v[1] for v in agg.values())
    print(f"{len(order)} launches, {tot / 1e3:.3f} ms total device time (cold-cache, serialised: compare shares)")
    print(f"{'time_us':>12} {'count':>6} {'share':>6}  kernel")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1]:12.1f} {v[0]:6d} {100 * v[1] / tot:5.1f}%  {k}")
    for o in order[:nfirst]:
        print(o)


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
