#!/bin/bash
# Round-end evidence on the final build: launch list of the default command, then the two bench arms.
TAG=${1:-r1s3}
# ncu replays every kernel: keep the pre-folds short (the captures below then sit at folds 38..41 of the run, where T still
# fits nine 15-bit windows; bench.py itself measures around fold 360, see PREFOLD there)
export VIMZ_BENCH_PREFOLD=32
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --msm-log2 20 > gpurun_out/${TAG}_launches_bench.log 2>&1
W=$(python tools/step_window.py gpurun_out/${TAG}_launches.csv 38 41)
echo "window $W"
python tools/launch_table.py gpurun_out/${TAG}_launches.csv $W > gpurun_out/${TAG}_launches_fold_step.txt
cat gpurun_out/${TAG}_launches_fold_step.txt
env -u VIMZ_BENCH_PREFOLD python bench.py > gpurun_out/${TAG}_bench_final.json 2> gpurun_out/${TAG}_bench_final.err
env -u VIMZ_BENCH_PREFOLD python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref_final.json 2>> gpurun_out/${TAG}_bench_final.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench_final.json").read().strip().splitlines()[-1])
r=json.loads(open("gpurun_out/${TAG}_bench_ref_final.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"]["value"], "ref", r["value"], "launches", d["gpu_launches"])
print("roofline", d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline"]["frac_of_peak_at_max_clock"], d["roofline"]["share_of_step"], d["roofline"]["launch_us_avg"])
print("hbm", d["roofline_hbm"]["achieved"], d["roofline_hbm"]["frac"])
print("msm", d["msm"][0]["mpts_per_s"], d["msm"][0]["accumulate_frac_of_imad_peak"])
print("clocks", d["clocks"])
PY
