#!/bin/bash
# quick check: subset of GPU parity tests + timings
mkdir -p gpurun_out
R=${1:-q}
timeout 900 python -m pytest tests -q -m gpu -k "golden and (leo7 or lti3 or lti4_2pnoise or syn5 or leo5)" 2>&1 | tail -5
timeout 600 python tools/time_scenario.py leo7 2 2>&1 | grep -v "rep 0" | tail -45
timeout 300 python tools/time_scenario.py lti3 2 2>&1 | tail -28
