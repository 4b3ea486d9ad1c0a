#!/bin/bash
run() { echo "== $*"; env "$@" python tools/host_breakdown.py 2>&1 | tail -3 | head -2 | cut -c1-330; }
run VIMZ_OPTS_PRIMARY=msm_seg_min_aux=4
run VIMZ_OPTS_PRIMARY=msm_seg_min_aux=5
run VIMZ_OPTS_PRIMARY=msm_seg_min_aux=6
run VIMZ_OPTS_PRIMARY=msm_seg_min_aux=12
