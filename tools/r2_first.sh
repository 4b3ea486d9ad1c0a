#!/bin/bash
# Round 2, first GPU call: the parity suite on the new tests, a baseline bench line, and the launch list of the configuration
# bench.py actually times (PREFOLD 260: folds 263.., tenth window of T in use, giant bucket present).
TAG=${1:-r2a}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
tail -5 gpurun_out/${TAG}_pytest.log
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 600 gpurun_out/${TAG}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --msm-log2 > gpurun_out/${TAG}_launches_bench.log 2>&1
W=$(python tools/step_window.py gpurun_out/${TAG}_launches.csv 266 269)
echo "window $W"
python tools/launch_table.py gpurun_out/${TAG}_launches.csv $W > gpurun_out/${TAG}_launches_fold_step.txt
cat gpurun_out/${TAG}_launches_fold_step.txt
# keep only the window of the big csv (gpurun_out is limited to 64 MiB)
python - <<PY
import sys
lo, hi = map(int, "$W".split())
out = []
for l in open("gpurun_out/${TAG}_launches.csv"):
    if l.startswith("==") : continue
    f = l.split(",", 1)[0].strip('"')
    if not f.isdigit() or lo <= int(f) <= hi: out.append(l)
open("gpurun_out/${TAG}_launches_window.csv", "w").writelines(out)
PY
rm -f gpurun_out/${TAG}_launches.csv
