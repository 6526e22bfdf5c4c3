#!/bin/bash
# quick check: subset of GPU parity tests + timings + ncu captures of the secondary kernels
mkdir -p gpurun_out
R=${1:-q}
timeout 900 python -m pytest tests -q -m gpu -k "golden and (leo7 or lti3 or lti4_2pnoise or syn5 or leo5)" 2>&1 | tail -5
timeout 600 python tools/time_scenario.py leo7 2 2>&1 | tail -15
timeout 300 python tools/time_scenario.py lti3 2 2>&1 | tail -13
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:KTpDce2 -s 2 -c 1 -o gpurun_out/prof_tpdce_$R -f python tools/profile_pass.py leo7 10 > gpurun_out/ncu_tpdce_$R.log 2>&1
tail -2 gpurun_out/ncu_tpdce_$R.log
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:KMsmtUpdate -s 23 -c 8 -o gpurun_out/prof_mu_$R -f python tools/profile_pass.py leo7 12 > gpurun_out/ncu_mu_$R.log 2>&1
tail -2 gpurun_out/ncu_mu_$R.log
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:KRegroup -s 9 -c 1 -o gpurun_out/prof_regroup_$R -f python tools/profile_pass.py leo7 11 > gpurun_out/ncu_regroup_$R.log 2>&1
tail -2 gpurun_out/ncu_regroup_$R.log
