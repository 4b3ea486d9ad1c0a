#!/bin/bash
python bench.py > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "bench rc $?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2g_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "serial", d.get("value_serial_calls"), "e2e", d["e2e"], "cpu", d["cpu_baseline"]["value"], "launches", d["gpu_launches"], "parity", d["parity_check"])
print("roofline", d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline"]["per_launch_class"])
print("msm", [(m["log2_points"], m["mpts_per_s"]) for m in d["msm"]])
print("clocks", d["clocks"], d.get("clock_verdict"))
print("configs", [(c["circuit"], round(c["steps_per_s"])) for c in d["configs"]["circuits"]], [(m["log2_points"], round(m["mpts_per_s"])) for m in d["configs"]["msm"]])
PY
