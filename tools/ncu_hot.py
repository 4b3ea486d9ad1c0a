#!/usr/bin/env python3
"""Hot spots of one kernel from `ncu -i X.ncu-rep --page source --csv [--print-source sass]`: bins the warp-stall
samples and executed instructions over address ranges (and prints the stall mix) so that a 50 KB kernel can be read.
usage: ncu -i rep --page source --csv | python tools/ncu_hot.py [kernel-substring] [bin-instructions]"""
import csv
import sys

want = sys.argv[1] if len(sys.argv) > 1 else ""
binsz = int(sys.argv[2]) if len(sys.argv) > 2 else 64
rows = list(csv.reader(sys.stdin))
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        name = rows[i][1]
        hdr = rows[i + 1]
        j = i + 2
        body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            if len(rows[j]) == len(hdr):
                body.append(rows[j])
            j += 1
        i = j
        if want not in name:
            continue
        col = {h: k for k, h in enumerate(hdr)}
        stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        tot_s = sum(int(r[col["# Samples"]] or 0) for r in body)
        tot_i = sum(int(r[col["Instructions Executed"]] or 0) for r in body)
        print(f"== {name[:100]}: {len(body)} instructions ({len(body) * 16 / 1024:.1f} KB), {tot_s} samples, {tot_i} warp-instructions executed")
        mix = {h: sum(int(r[col[h]] or 0) for r in body) for h in stalls}
        print("   stall mix:", ", ".join(f"{h[6:]} {100 * v / max(tot_s, 1):.0f}%" for h, v in sorted(mix.items(), key=lambda kv: -kv[1])[:8]))
        for b in range(0, len(body), binsz):
            chunk = body[b:b + binsz]
            s = sum(int(r[col["# Samples"]] or 0) for r in chunk)
            ins = sum(int(r[col["Instructions Executed"]] or 0) for r in chunk)
            if s * 50 < tot_s and ins * 50 < tot_i:
                continue
            top = max(chunk, key=lambda r: int(r[col["# Samples"]] or 0))
            cm = {h: sum(int(r[col[h]] or 0) for r in chunk) for h in stalls}
            lead = ", ".join(f"{h[6:]} {v}" for h, v in sorted(cm.items(), key=lambda kv: -kv[1])[:3] if v)
            print(f"   [{b:5d}..{b + len(chunk):5d}) samples {100 * s / max(tot_s, 1):5.1f}%  executed {100 * ins / max(tot_i, 1):5.1f}%  {lead:48s} hottest: {top[col['Source']].strip()[:60]}")
    else:
        i += 1
