#!/usr/bin/env python3
"""IDs of the kernels of fold steps [first_step, last_step] (1-based, both curves) inside an ncu launch list of bench.py:
every step launches k_matvec_stream twice (secondary curve, then primary), so the (2*(s-1)+1)-th such launch opens step s.
usage: python tools/step_window.py launches.csv first_step last_step   ->   "first_id last_id" """
import sys

sys.path.insert(0, __import__("os").path.dirname(__file__))
from launch_table import load  # noqa: E402

rows = load(sys.argv[1])
first, last = int(sys.argv[2]), int(sys.argv[3])
marks = [r[0] for r in rows if r[1] == "k_matvec_stream"]
lo = marks[2 * (first - 1)]
hi = marks[2 * last] - 1 if 2 * last < len(marks) else rows[-1][0]
# kernels of the aux lane (the W2 digit pass) may be listed just before the step's first cross term: pull them in
ids = [r[0] for r in rows]
i = ids.index(lo)
while i > 0 and rows[i - 1][1] in ("k_msm_digits",) and lo - rows[i - 1][0] <= 2:
    i -= 1
    lo = rows[i][0]
print(lo, hi)
