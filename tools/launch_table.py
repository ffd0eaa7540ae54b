#!/usr/bin/env python3
"""Markdown table (launches, mean us, share of GPU time) from an ncu launch list:
   ncu --metrics gpu__time_duration.sum --clock-control none -c N --csv --log-file launches.csv <command>"""
import csv, sys
from collections import OrderedDict


def main():
    rows = []
    with open(sys.argv[1]) as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.DictReader(lines)
    agg = OrderedDict()
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
        a = agg.setdefault(r["Kernel Name"][:70], [0, 0.0])
        a[0] += 1; a[1] += us
    tot = sum(a[1] for a in agg.values())
    print("| kernel | launches | mean µs | share of GPU time |")
    print("|---|---|---|---|")
    for k, (n, t) in agg.items():
        print(f"| `{k}` | {n} | {t / n:.1f} | {t / tot:.3f} |")
    print(f"\ntotal GPU time in the capture: {tot / 1000:.2f} ms over {sum(a[0] for a in agg.values())} launches")


if __name__ == "__main__":
    main()
