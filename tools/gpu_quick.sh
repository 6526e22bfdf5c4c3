#!/bin/bash
# quick check: subset of GPU parity tests + timings + one full ncu capture of KGTable
mkdir -p gpurun_out
R=${1:-q}
timeout 900 python -m pytest tests -q -m gpu -k "golden and (leo7 or lti3 or lti4_2pnoise or syn5)" 2>&1 | tail -5
timeout 600 python tools/time_scenario.py leo7 2 2>&1 | tail -15
timeout 300 python tools/time_scenario.py lti3 2 2>&1 | tail -13
timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:KGTable -s 60 -c 1 -o gpurun_out/prof_gtable_$R -f python tools/profile_pass.py leo7 11 > gpurun_out/ncu_full_$R.log 2>&1
tail -3 gpurun_out/ncu_full_$R.log
