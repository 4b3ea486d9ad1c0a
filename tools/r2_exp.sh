#!/bin/bash
run() { echo "== $*"; env "$@" python tools/host_breakdown.py 2>&1 | tail -3 | head -1; }
run VIMZ_DIRECT_BPS=3
run VIMZ_DIRECT_BPS=2
run VIMZ_DIRECT_BPS=4
python tools/timeline.py 260 2>/dev/null | tail -36 | head -14
