#!/bin/bash
for c in 13 14 10; do echo "direct_c=$c"; VIMZ_DIRECT_C=$c python tools/host_breakdown.py 2>&1 | tail -3 | head -1; done
