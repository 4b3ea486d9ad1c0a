#!/bin/bash
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
python bench.py --steps 200 --warmup 5 > gpurun_out/ks8_bench.json 2> gpurun_out/ks8_bench.err || tail -5 gpurun_out/ks8_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/ks8_bench.json").read().strip().splitlines()[-1])
e = d["e2e"]
print(round(d["value"], 1), "steps/s (serial calls", round(d["value_serial_calls"], 1), ") e2e", round(e["value"], 1), "(serial", round(e["serial_calls_value"], 1), ") plain", round(e["plain_call_value"],1), "pageable", round(e["pageable_value"],1), "roofline", round(d["roofline"]["frac"], 3), "parity", (d.get("parity_check") or {}).get("equal"), "cpu", d["cpu_baseline"]["value"])
print(d["config"].get("secondary_key"))
print("msm", [(m["log2_points"], round(m["mpts_per_s"],1)) for m in d.get("msm", [])])
c = d.get("configs") or {}
print("configs keys", list(c.keys()))
for k, v in (c.get("circuits") or {}).items(): print(k, {kk: (round(vv,1) if isinstance(vv,float) else vv) for kk, vv in v.items() if kk in ("steps_per_s","e2e_steps_per_s","error")})
print("msm sweep", [(m["log2_points"], round(m["mpts_per_s"],1), m.get("result_equals_closed_form")) for m in (c.get("msm_sweep") or [])])
PY
tail -3 gpurun_out/ks8_bench.err
