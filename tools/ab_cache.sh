#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -12
run() { local label=$1; shift; env "$@" python bench.py --steps 100 --warmup 5 --no-cpu-baseline --msm-log2 > /tmp/ab.json 2>/tmp/ab.err || { echo "$label FAILED"; tail -3 /tmp/ab.err; return; }
  python - "$label" <<'PY'
import json, sys
d = json.loads(open("/tmp/ab.json").read().strip().splitlines()[-1])
f = lambda ph: {k: round(v["ms_per_step"], 3) for k, v in d[ph].items() if v["calls"]}
print(sys.argv[1], round(d["value"], 1), "steps/s e2e", round(d["e2e"]["value"], 1), "prim", f("phases_primary"), "sec", f("phases_secondary"), flush=True)
PY
}
run "no cache" VIMZ_CROSS_CACHE=0
run "cache minb2"
run "cache minb3" VIMZ_GPU_LIB=$PWD/build/variants/minb3.so
run "cache minb4" VIMZ_GPU_LIB=$PWD/build/variants/minb4.so
