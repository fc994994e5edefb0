#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name: count, mean and min duration.

    python tools/ncu_launch_table.py launches.csv
"""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    head = rows[h]
    ix = {k: i for i, k in enumerate(head)}
    agg = collections.defaultdict(list)
    for r in rows[h + 1:]:
        if len(r) >= len(head) and r[ix["Metric Name"]] == "gpu__time_duration.sum":
            agg[r[ix["Kernel Name"]][:72]].append(float(r[ix["Metric Value"]].replace(",", "")))
    for k, v in agg.items():
        print(f"{k:74s} n={len(v):4d} mean={sum(v) / len(v) / 1e3:8.2f} us  min={min(v) / 1e3:8.2f} us")


if __name__ == "__main__":
    main(sys.argv[1])
