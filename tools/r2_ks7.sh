#!/bin/bash
python -m pytest tests/test_gpu_r1cs.py -m gpu -x -q -k "staged or booleanity" > gpurun_out/ks7_pytest.log 2>&1; tail -2 gpurun_out/ks7_pytest.log
python bench.py --steps 200 --warmup 5 --no-configs > gpurun_out/ks7_bench.json 2> gpurun_out/ks7_bench.err || tail -5 gpurun_out/ks7_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/ks7_bench.json").read().strip().splitlines()[-1])
e = d["e2e"]
print(round(d["value"], 1), "steps/s (serial calls", round(d["value_serial_calls"], 1), ") e2e", round(e["value"], 1), "(serial", round(e["serial_calls_value"], 1), ") plain", round(e["plain_call_value"],1), "pageable", round(e["pageable_value"],1), "roofline", round(d["roofline"]["frac"], 3), "parity", (d.get("parity_check") or {}).get("equal"), "cpu", d["cpu_baseline"]["value"])
print(d["config"].get("secondary_key"))
print("msm", [(m["log2_points"], round(m["mpts_per_s"],1)) for m in d.get("msm", [])])
PY
