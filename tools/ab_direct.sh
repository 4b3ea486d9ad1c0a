#!/bin/bash
# A/B of the direct-table MSM path on the fold step (secondary curve): blocks per SM, inlined vs called madd, bucket baseline
run() {  # label, env...
  local label=$1; shift
  env "$@" python bench.py --steps 60 --warmup 5 --no-cpu-baseline --msm-log2 > /tmp/ab.json 2>/tmp/ab.err || { echo "$label FAILED"; tail -3 /tmp/ab.err; return; }
  python - "$label" <<'PY'
import json, sys
d = json.loads(open("/tmp/ab.json").read().strip().splitlines()[-1])
f = lambda ph: {k: round(v["ms_per_step"], 3) for k, v in d[ph].items() if v["calls"]}
print(sys.argv[1], round(d["value"], 1), "steps/s e2e", round(d["e2e"]["value"], 1), "sec", f("phases_secondary"), flush=True)
PY
}
run "buckets(direct_max=0)" VIMZ_DIRECT_MAX=0
run "direct call bps=4" VIMZ_DIRECT_BPS=4
run "direct call bps=2" VIMZ_DIRECT_BPS=2
run "direct call bps=1" VIMZ_DIRECT_BPS=1
run "direct inline bps=4" VIMZ_GPU_LIB=$PWD/build/variants/direct_inline.so VIMZ_DIRECT_BPS=4
run "direct inline bps=2" VIMZ_GPU_LIB=$PWD/build/variants/direct_inline.so VIMZ_DIRECT_BPS=2
ncu --set full --clock-control none --import-source on -k regex:k_msm_direct -s 40 -c 2 -o gpurun_out/s3_direct python bench.py --steps 4 --warmup 3 --no-cpu-baseline --msm-log2 > /dev/null 2>&1
ncu -i gpurun_out/s3_direct.ncu-rep --page raw --csv > gpurun_out/s3_direct_raw.csv 2>/dev/null
ls -la gpurun_out/ | head
