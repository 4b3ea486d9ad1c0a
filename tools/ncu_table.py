#!/usr/bin/env python3
"""One line per captured launch from `ncu -i X.ncu-rep --page raw --csv` files: duration, registers, DRAM bytes, L2 hit rate,
resident warps, issue-slot and multiplier-pipe activity, and the three largest warp-stall reasons.
usage: python tools/ncu_table.py raw1.csv [raw2.csv ...] > profiles/r2_ncu_table.txt"""
import csv
import re
import sys

print(f"{'kernel':44s} {'grid':>12s} {'us':>8s} {'regs':>5s} {'dramRdMB':>9s} {'dramWrMB':>9s} {'L2hit%':>7s} {'warps%':>7s} {'issue%':>7s} {'fmaH%':>6s}  top stalls (share of sampled warp-cycles)")
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h]
    if not stall_cols:
        stall_cols = [h for h in hdr if h.startswith("smsp__average_warp_latency_issue_stalled_") and h.endswith(".ratio")]

    def f(v, key, scale=1.0):
        try:
            x = float(v[col[key]].replace(",", ""))
        except Exception:
            return float("nan")
        u = units[col[key]]
        if u == "Mbyte": x *= 1.0
        elif u == "Kbyte": x /= 1e3
        elif u == "byte": x /= 1e6
        elif u == "Gbyte": x *= 1e3
        elif u == "ms": x *= 1e3
        elif u == "ns": x /= 1e3
        return x * scale

    for v in rows[2:]:
        name = re.sub(r"^void ", "", v[col["Kernel Name"]])
        name = re.sub(r"\(.*", "", name)
        stalls = []
        for h in stall_cols:
            try:
                stalls.append((float(v[col[h]].replace(",", "")), re.sub(r"smsp__average_warps?_(latency_)?issue_stalled_|_per_issue_active\.ratio|\.ratio", "", h)))
            except Exception:
                pass
        tot = sum(s for s, _ in stalls) or 1.0
        top = ", ".join(f"{n} {100 * s / tot:.0f}%" for s, n in sorted(stalls, reverse=True)[:3])
        print(f"{name[:44]:44s} {v[col['Grid Size']].replace(' ', ''):>12s} {f(v, 'gpu__time_duration.sum'):8.1f} {v[col['launch__registers_per_thread']]:>5s} "
              f"{f(v, 'dram__bytes_read.sum'):9.2f} {f(v, 'dram__bytes_write.sum'):9.2f} {f(v, 'lts__t_sector_hit_rate.pct'):7.1f} "
              f"{f(v, 'sm__warps_active.avg.pct_of_peak_sustained_active'):7.1f} {f(v, 'sm__issue_active.avg.pct_of_peak_sustained_elapsed'):7.1f} "
              f"{f(v, 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed'):6.1f}  {top}")
