#!/bin/bash
# booleanity-row fold (K_S): parity tests, then A/B of the fold-step bench with the option on / off, then the whole GPU suite
python -m pytest tests/test_gpu_r1cs.py -m gpu -x -q -k "booleanity or full_size" > gpurun_out/ks_pytest_new.log 2>&1; tail -3 gpurun_out/ks_pytest_new.log
for o in 1 0; do
  VIMZ_OPTS=bitrow_fold=$o python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-configs --msm-log2 > gpurun_out/ks_bench_$o.json 2> gpurun_out/ks_bench_$o.err || tail -5 gpurun_out/ks_bench_$o.err
  python - $o <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/ks_bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
f = lambda ph: {k: round(v["ms_per_step"], 3) for k, v in d[ph].items() if v["calls"]}
print("bitrow_fold", sys.argv[1], round(d["value"], 1), "steps/s e2e", round(d["e2e"]["value"], 1), "roofline", round(d["roofline"]["frac"], 3), d["roofline"]["insertions_per_launch"], d["roofline"]["launch_us_avg"], "parity", d.get("parity_check", {}).get("equal"))
print(" prim", f("phases_primary")); print(" sec", f("phases_secondary"))
PY
done
VIMZ_OPTS=bitrow_fold=1 python tools/host_breakdown.py 2>&1 | tail -3
python -m pytest tests -m gpu -x -q > gpurun_out/ks_pytest_all.log 2>&1; tail -3 gpurun_out/ks_pytest_all.log
