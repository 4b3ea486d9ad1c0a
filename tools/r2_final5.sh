#!/bin/bash
# Very last evidence of round 2: default bench, timeline and the tail kernel's counters on the final geometry (16 blocks per sum)
python bench.py > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; echo "bench rc $?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2j_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"]["value"], "launches", d["gpu_launches"], "parity", d["parity_check"]["equal"])
print("roofline", d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline"]["per_launch_class"])
print("msm", [(m["log2_points"], m["mpts_per_s"], m["accumulate_frac_of_imad_peak"]) for m in d["msm"]])
print("configs", [(c["circuit"], round(c["steps_per_s"])) for c in d["configs"]["circuits"]], [(m["log2_points"], round(m["mpts_per_s"])) for m in d["configs"]["msm"]])
print("clocks", d["clocks"], d.get("clock_verdict"))
PY
python tools/timeline.py 260 > gpurun_out/r2j_timeline_fold_step.txt 2>/dev/null
timeout 150 ncu --set full --clock-control none -k regex:k_reduce_tail -s 528 -c 2 -o gpurun_out/r2j_tail python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-configs --msm-log2 > /dev/null 2>&1
ncu -i gpurun_out/r2j_tail.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summary.py > gpurun_out/r2j_ncu_reduce_tail.txt 2>/dev/null
rm -f gpurun_out/r2j_tail.ncu-rep
grep -n "gpu__time_duration.sum\|launch__grid_size" gpurun_out/r2j_ncu_reduce_tail.txt | head -4
