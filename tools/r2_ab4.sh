#!/bin/bash
# A/B: dedicated squaring (-DVIMZ_FP_SQR=1, build/variants/sqr.so) against the default library; correctness first
python -m pytest tests/test_gpu_field.py tests/test_gpu_msm.py -m gpu -x -q 2>&1 | tail -2
python -m pytest tests/test_gpu_field.py -m gpu -x -q 2>&1 | tail -1
run() {
  echo "== $1"
  for r in 1 2; do VIMZ_GPU_LIB=$PWD/$1 python tools/host_breakdown.py 2>&1 | tail -3 | head -1; done
  VIMZ_GPU_LIB=$PWD/$1 python bench.py --msm-only --msm-log2 16 20 24 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
for m in d['msm']: print('  msm', m['log2_points'], round(m['mpts_per_s'],1), 'Mpts/s acc_ms', round(m['accumulate_ms'],3), 'frac', round(m['accumulate_frac_of_imad_peak'],3), m['result_equals_closed_form'])
"
}
run vimz_b200/libvimz_gpu.so
run build/variants/noopaque.so
python -m pytest tests/test_gpu_r1cs.py tests/test_gpu_msm.py -m gpu -x -q 2>&1 | tail -1
