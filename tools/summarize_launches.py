#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (count, total, mean, share)."""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"])[:70]
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else v * 1e3 if unit == "ms" else v
        agg.setdefault(name, []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print(f"{'kernel':72s} {'n':>5s} {'sum_us':>10s} {'mean_us':>9s} {'min_us':>8s} {'max_us':>8s} {'share':>6s}")
    for k, v in agg.items():
        print(f"{k:72s} {len(v):5d} {sum(v):10.1f} {sum(v) / len(v):9.1f} {min(v):8.1f} {max(v):8.1f} {100 * sum(v) / tot:5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
