#!/bin/bash
for c in brightness crop; do
python bench.py --circuit $c --steps 100 --no-cpu-baseline --no-configs --msm-log2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$c', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'plain', round(d['e2e']['plain_call_value'],1), d['config']['msm_window_bits'])"
VIMZ_BENCH_PREFOLD=8 python bench.py --circuit $c --steps 10 --no-cpu-baseline --no-configs --msm-log2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$c prefold8 steps10', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'plain', round(d['e2e']['plain_call_value'],1), d['config']['msm_window_bits'])"
done
