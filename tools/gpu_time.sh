#!/bin/bash
# timing table only (last repetition, phase timing off)
mkdir -p gpurun_out
timeout 600 python tools/time_scenario.py ${1:-leo7} 3 > gpurun_out/time_${1:-leo7}.log 2>&1; tail -8 gpurun_out/time_${1:-leo7}.log | cut -c1-200
