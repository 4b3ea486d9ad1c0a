#!/bin/bash
# A/B: accumulation kernel compiled for 5 / 6 resident blocks per SM (96 / 80 registers, spills) against the default 4 (128 registers)
run() { # lib blocks
  echo "== $1 blocks=$2"
  VIMZ_GPU_LIB=$PWD/$1 VIMZ_ACC_BLOCKS=$2 python tools/host_breakdown.py 2>&1 | tail -3 | head -1
  VIMZ_GPU_LIB=$PWD/$1 VIMZ_ACC_BLOCKS=$2 python bench.py --msm-only --msm-log2 20 24 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
for m in d['msm']: print('  msm', m['log2_points'], round(m['mpts_per_s'],1), 'Mpts/s acc_ms', round(m['accumulate_ms'],3), 'frac', round(m['accumulate_frac_of_imad_peak'],3), m['result_equals_closed_form'])
"
}
run vimz_b200/libvimz_gpu.so 4
run build/variants/acc5.so 5
run build/variants/acc5.so 4
run build/variants/acc6.so 6
run build/variants/acc6.so 5
