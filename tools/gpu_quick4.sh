#!/bin/bash
# quick check: subset of GPU parity tests + the timing table of the last repetition (phase timing off)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -k "golden and (leo7 or lti3 or lti4_2pnoise or syn5 or leo5)" 2>&1 | tail -5
timeout 600 python tools/time_scenario.py leo7 3 > gpurun_out/time_leo7.log 2>&1; tail -15 gpurun_out/time_leo7.log | cut -c1-200
timeout 300 python tools/time_scenario.py lti3 2 > gpurun_out/time_lti3.log 2>&1; tail -14 gpurun_out/time_lti3.log | cut -c1-200
