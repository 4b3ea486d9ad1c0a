#!/bin/bash
# fold steps/s vs window size of the primary key on the current build, timed where bench.py times (PREFOLD 260)
for c in 14 15 16 17; do
  VIMZ_WINDOW_PALLAS=$c python bench.py --steps 100 --no-cpu-baseline --msm-log2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
ph=d['phases_primary']
print('c=$c', round(d['value'],1), 'steps/s', round(d['ms_per_step'],4), 'ms; e2e', round(d['e2e']['value'],1), {k: round(v['ms_per_step'],3) for k,v in ph.items()}, 'entries/launch', d['roofline']['insertions_per_launch'])
"
done
