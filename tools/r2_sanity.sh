#!/bin/bash
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python -m pytest tests/test_gpu_msm.py tests/test_gpu_r1cs.py -m gpu -x -q 2>&1 | tail -1
python bench.py --steps 100 --warmup 5 --no-configs > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; echo "bench rc $?"
python -c "
import json
d=json.loads(open('gpurun_out/r2i_bench.json').read().strip().splitlines()[-1])
print('value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'parity', d['parity_check']['equal'], 'msm', [(m['log2_points'], round(m['mpts_per_s'],1)) for m in d['msm']], 'frac', round(d['roofline']['frac'],3), d['roofline']['traffic_source'][:40])
"
