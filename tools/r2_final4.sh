#!/bin/bash
# Last evidence pass of round 2 (final build: paired chunk sums + dedicated squaring): the r2_profile.sh captures as r2h_*,
# then the full GPU suite and the default bench.
bash tools/r2_profile.sh r2h > gpurun_out/r2h_call.log 2>&1
tail -30 gpurun_out/r2h_call.log | head -12
python -m pytest tests -m gpu -x -q > gpurun_out/r2h_pytest.log 2>&1; tail -2 gpurun_out/r2h_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench rc $?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2h_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"]["value"], "launches", d["gpu_launches"], "parity", d["parity_check"]["equal"])
print("roofline", d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline"]["per_launch_class"])
print("msm", [(m["log2_points"], m["mpts_per_s"], m["accumulate_frac_of_imad_peak"]) for m in d["msm"]])
print("clocks", d["clocks"], d.get("clock_verdict"))
PY
