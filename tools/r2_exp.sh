#!/bin/bash
run() { echo "== $*"; env "$@" python tools/host_breakdown.py 2>&1 | tail -3 | head -2; }
run VIMZ_X=1
run VIMZ_WINDOW_PALLAS=16
run VIMZ_GPU_LIB=build/variants/mid32.so
run VIMZ_GPU_LIB=build/variants/mid32.so VIMZ_WINDOW_PALLAS=16
run VIMZ_X=1
