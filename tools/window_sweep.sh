#!/bin/bash
# sweep the MSM window size: fold step (primary Pallas key 2^17, secondary Vesta key 2^14) and stand-alone 2^20 MSM
for wp in 13 14 15 16 17; do
  VIMZ_WINDOW_PALLAS=$wp python bench.py --steps 20 --warmup 3 --no-cpu-baseline --msm-log2 > /tmp/ws.json 2>/tmp/ws.err
  python - <<PY
import json
d=json.loads(open("/tmp/ws.json").read().strip().splitlines()[-1])
print("fold pallas c=$wp", round(d["value"],1), "steps/s", round(d["ms_per_step"],3), "ms", {k: round(v["ms_per_step"],3) for k,v in d["phases_primary"].items()})
PY
done
for ws in 10 11 12 13 14; do
  VIMZ_WINDOW_VESTA=$ws python bench.py --steps 20 --warmup 3 --no-cpu-baseline --msm-log2 > /tmp/ws.json 2>/tmp/ws.err
  python - <<PY
import json
d=json.loads(open("/tmp/ws.json").read().strip().splitlines()[-1])
print("fold vesta c=$ws", round(d["value"],1), "steps/s", round(d["ms_per_step"],3), "ms", {k: round(v["ms_per_step"],3) for k,v in d["phases_secondary"].items()})
PY
done
for w in 15 16 17 18 19 20; do
  VIMZ_WINDOW_MSM=$w python bench.py --steps 3 --warmup 3 --no-cpu-baseline --msm-only --msm-log2 20 > /tmp/ws.json 2>/tmp/ws.err
  python - <<PY
import json
d=json.loads(open("/tmp/ws.json").read().strip().splitlines()[-1])
for m in d["msm"]:
    print("msm 2^20 c=$w", {k: (round(v,3) if isinstance(v,float) else v) for k,v in m.items() if k in ("mpts_per_s","ms_per_msm","windows","accumulate_ms","sort_ms","reduce_ms","accumulate_frac_of_imad_peak")})
PY
done
