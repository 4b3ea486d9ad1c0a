#!/bin/bash
python -m pytest tests -m gpu -x -q > gpurun_out/ks6_pytest.log 2>&1; tail -2 gpurun_out/ks6_pytest.log
python bench.py --steps 200 --warmup 5 > gpurun_out/ks6_bench.json 2> gpurun_out/ks6_bench.err || tail -5 gpurun_out/ks6_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/ks6_bench.json").read().strip().splitlines()[-1])
print(round(d["value"], 1), "steps/s e2e", round(d["e2e"]["value"], 1), "plain", round(d["e2e"]["plain_call_value"],1), "pageable", round(d["e2e"]["pageable_value"],1), "roofline", round(d["roofline"]["frac"], 3), "parity", (d.get("parity_check") or {}).get("equal"), "cpu", d["cpu_baseline"]["value"])
print("msm", [(m["log2_points"], round(m["mpts_per_s"],1)) for m in d.get("msm", [])])
PY
