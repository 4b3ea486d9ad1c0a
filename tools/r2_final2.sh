#!/bin/bash
# Final-build validation (round 2, after the paired chunk sums): full GPU suite, smoke, both bench arms, crop timeline.
python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1; tail -2 gpurun_out/r2g_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python bench.py > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "bench rc $?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2g_bench_ref.json 2>> gpurun_out/r2g_bench.err; echo "ref rc $?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2g_bench.json").read().strip().splitlines()[-1])
r=json.loads(open("gpurun_out/r2g_bench_ref.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"]["value"], "ref", r["value"], "launches", d["gpu_launches"], "parity", d["parity_check"])
print("roofline", d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline"]["per_launch_class"])
print("msm", [(m["log2_points"], m["mpts_per_s"]) for m in d["msm"]])
print("clocks", d["clocks"])
PY
VIMZ_TL_CIRCUIT=crop python tools/timeline.py 30 > gpurun_out/r2g_timeline_crop.txt 2>/dev/null; tail -70 gpurun_out/r2g_timeline_crop.txt | head -64
python tools/timeline.py 260 > gpurun_out/r2g_timeline_fold_step.txt 2>/dev/null
python tools/host_breakdown.py 2>&1 | tail -3 | head -1
