#!/bin/bash
python -m pytest tests/test_gpu_r1cs.py tests/test_gpu_recursive.py -m gpu -x -q -k "booleanity or full_size or recursive or chain" > gpurun_out/ks2_pytest_new.log 2>&1; tail -3 gpurun_out/ks2_pytest_new.log
python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-configs --msm-log2 > gpurun_out/ks2_bench.json 2> gpurun_out/ks2_bench.err || tail -5 gpurun_out/ks2_bench.err
python - <<'PY'
import json, sys
d = json.loads(open("gpurun_out/ks2_bench.json").read().strip().splitlines()[-1])
f = lambda ph: {k: round(v["ms_per_step"], 3) for k, v in d[ph].items() if v["calls"]}
print(round(d["value"], 1), "steps/s e2e", round(d["e2e"]["value"], 1), "roofline", round(d["roofline"]["frac"], 3), "parity", (d.get("parity_check") or {}).get("equal"))
print(" prim", f("phases_primary")); print(" sec", f("phases_secondary"))
PY
python tools/host_breakdown.py 2>&1 | tail -3 | head -1
python tools/timeline.py 260 > gpurun_out/ks2_timeline.txt 2>/dev/null; tail -45 gpurun_out/ks2_timeline.txt
