#!/bin/bash
# Full ncu captures of named kernels on one scenario pass.  Usage: tools/gpu_ncu.sh TAG SCENARIO STEPS "REGEX:SKIP" ...
# Each REGEX:SKIP captures one launch (the SKIP-th match, 0-based) into gpurun_out/prof_<regex>_<TAG>.ncu-rep
mkdir -p gpurun_out
TAG=$1; SCEN=$2; STEPS=$3; shift 3
for spec in "$@"; do
  RX=${spec%%:*}; SKIP=${spec##*:}
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:$RX -s $SKIP -c 1 \
    -o gpurun_out/prof_${RX}_$TAG -f python tools/profile_pass.py $SCEN $STEPS > gpurun_out/ncu_${RX}_$TAG.log 2>&1
  tail -2 gpurun_out/ncu_${RX}_$TAG.log
done
