#!/bin/bash
python tools/host_breakdown.py 2>&1 | tail -3 | head -1
python tools/host_breakdown.py 2>&1 | tail -3 | head -1
