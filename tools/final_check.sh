python -m pytest tests -m gpu -x -q > gpurun_out/s3_pytest_final.log 2>&1; tail -2 gpurun_out/s3_pytest_final.log
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py > gpurun_out/s3_bench_final.json 2> gpurun_out/s3_bench_final.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s3_bench_ref_final.json 2>> gpurun_out/s3_bench_final.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/s3_bench_final.json").read().strip().splitlines()[-1])
r=json.loads(open("gpurun_out/s3_bench_ref_final.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"]["value"], "ref", r["value"], "launches", d["gpu_launches"])
print("roofline", d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline"]["frac_of_peak_at_max_clock"], d["roofline"]["share_of_step"])
print("msm", d["msm"])
print("clocks", d["clocks"])
PY
for dist in witness edge; do python bench.py --msm-only --msm-dist $dist --msm-log2 17 20 > gpurun_out/s3_msm_$dist.json 2>>gpurun_out/s3_bench_final.err; python -c "
import json
d=json.loads(open('gpurun_out/s3_msm_$dist.json').read().strip().splitlines()[-1])
for m in d['msm']: print('$dist', m['log2_points'], round(m['mpts_per_s'],1), 'Mpts/s', round(m['ms_per_msm'],3), 'ms', m['window_bits'])
"; done
python bench.py --msm-only --msm-log2 16 18 20 22 24 > gpurun_out/s3_msm_sweep.json 2>>gpurun_out/s3_bench_final.err; python -c "
import json
d=json.loads(open('gpurun_out/s3_msm_sweep.json').read().strip().splitlines()[-1])
for m in d['msm']: print('uniform', m['log2_points'], round(m['mpts_per_s'],1), 'Mpts/s', round(m['ms_per_msm'],3), 'ms c=', m['window_bits'], round(m.get('accumulate_frac_of_imad_peak',0),3))
"
