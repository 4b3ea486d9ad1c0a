#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / mean / total and share.
usage: python tools/launch_table.py gpurun_out/launches.csv [first_id last_id]"""
import csv
import re
import sys
from collections import OrderedDict


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = []
    for x in csv.DictReader(lines):
        if x.get("Metric Name") == "gpu__time_duration.sum":
            name = x["Kernel Name"]
            curve = "Pallas" if "Pallas" in name else "Vesta" if "Vesta" in name else "Bn" if "Bn" in name else "Grumpkin" if "Grumpkin" in name else ""
            short = re.sub(r"^void ", "", name)
            short = re.sub(r"[<(].*", "", short.replace("vimz::", ""))
            rows.append((int(x["ID"]), short, curve, float(x["Metric Value"].replace(",", "")) / 1e3, x["Grid Size"], x["Block Size"]))
    return rows


def main():
    rows = load(sys.argv[1])
    if len(sys.argv) >= 4:
        lo, hi = int(sys.argv[2]), int(sys.argv[3])
        rows = [r for r in rows if lo <= r[0] <= hi]
    agg = OrderedDict()
    for _, short, curve, us, grid, block in rows:
        k = (short, curve)
        a = agg.setdefault(k, [0, 0.0, grid, block])
        a[0] += 1
        a[1] += us
    total = sum(a[1] for a in agg.values())
    print(f"{'kernel':34s} {'field/curve':9s} {'launches':>8s} {'mean us':>10s} {'total us':>11s} {'share':>7s}  grid block")
    for (short, curve), (n, us, grid, block) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{short:34s} {curve:9s} {n:8d} {us / n:10.1f} {us:11.1f} {100 * us / total:6.1f}%  {grid} {block}")
    print(f"{'TOTAL':34s} {'':9s} {sum(a[0] for a in agg.values()):8d} {'':10s} {total:11.1f}")


if __name__ == "__main__":
    main()
