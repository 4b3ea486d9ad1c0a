#!/bin/bash
# A/B: fold-step bench with the default library and every build/variants/*.so (VIMZ_GPU_LIB override)
for lib in vimz_b200/libvimz_gpu.so build/variants/*.so; do
  [ -f "$lib" ] || continue
  VIMZ_GPU_LIB=$PWD/$lib python bench.py --steps 100 --warmup 5 --no-cpu-baseline --msm-log2 > /tmp/ab.json 2>/tmp/ab.err || { echo "$lib FAILED"; tail -3 /tmp/ab.err; continue; }
  python - "$lib" <<'PY'
import json, sys
d = json.loads(open("/tmp/ab.json").read().strip().splitlines()[-1])
f = lambda ph: {k: round(v["ms_per_step"], 3) for k, v in d[ph].items() if v["calls"]}
print(sys.argv[1], round(d["value"], 1), "steps/s e2e", round(d["e2e"]["value"], 1), "prim", f("phases_primary"), "sec", f("phases_secondary"))
PY
done
