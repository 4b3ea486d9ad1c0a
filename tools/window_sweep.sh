#!/bin/bash
# sweep the MSM window size of the primary (Pallas) and secondary (Vesta) commitment keys for the fold step
for wp in 13 14 15 16 17; do
  VIMZ_WINDOW_PALLAS=$wp python bench.py --steps 20 --warmup 3 --no-cpu-baseline --msm-log2 > /tmp/ws.json 2>/tmp/ws.err
  python - <<PY
import json
d=json.loads(open("/tmp/ws.json").read().strip().splitlines()[-1])
print("pallas c=$wp", round(d["value"],1), "steps/s", round(d["ms_per_step"],3), "ms", {k: round(v["ms_per_step"],3) for k,v in d["phases_primary"].items()})
PY
done
for ws in 10 11 12 13 14 15; do
  VIMZ_WINDOW_VESTA=$ws python bench.py --steps 20 --warmup 3 --no-cpu-baseline --msm-log2 > /tmp/ws.json 2>/tmp/ws.err
  python - <<PY
import json
d=json.loads(open("/tmp/ws.json").read().strip().splitlines()[-1])
print("vesta c=$ws", round(d["value"],1), "steps/s", round(d["ms_per_step"],3), "ms", {k: round(v["ms_per_step"],3) for k,v in d["phases_secondary"].items()})
PY
done
