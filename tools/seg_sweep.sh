#!/bin/bash
# sweep the shortest accumulation segment (option msm_seg_min) for the two fold-step curves
for sp in 4 6 8 12; do
  VIMZ_SEG_MIN_PALLAS=$sp python bench.py --steps 30 --warmup 3 --no-cpu-baseline --msm-log2 > /tmp/ws.json 2>/tmp/ws.err
  python - <<PY
import json
d=json.loads(open("/tmp/ws.json").read().strip().splitlines()[-1])
print("fold pallas seg_min=$sp", round(d["value"],1), "steps/s", {k: round(v["ms_per_step"],3) for k,v in d["phases_primary"].items()})
PY
done
for ss in 2 3 4 6 8; do
  VIMZ_SEG_MIN_VESTA=$ss python bench.py --steps 30 --warmup 3 --no-cpu-baseline --msm-log2 > /tmp/ws.json 2>/tmp/ws.err
  python - <<PY
import json
d=json.loads(open("/tmp/ws.json").read().strip().splitlines()[-1])
print("fold vesta seg_min=$ss", round(d["value"],1), "steps/s", {k: round(v["ms_per_step"],3) for k,v in d["phases_secondary"].items()})
PY
done
