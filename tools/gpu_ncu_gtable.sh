#!/bin/bash
# One full ncu capture (with source) of the last KGTable launch of MU 11 of the LEO7 window. Output -> gpurun_out/
mkdir -p gpurun_out
R=${1:-r01}
timeout 300 python tools/profile_pass.py leo7 > gpurun_out/profile_pass_$R.log 2>&1
tail -13 gpurun_out/profile_pass_$R.log
SKIP=$(python - <<PY
import re
c=[int(m.group(1)) for m in re.finditer(r"cumulative (\d+)", open("gpurun_out/profile_pass_$R.log").read())]
print(c[10]-1)
PY
)
echo "skip $SKIP"
timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:KGTable -s $SKIP -c 1 -o gpurun_out/prof_gtable_$R -f python tools/profile_pass.py leo7 11 > gpurun_out/ncu_full_$R.log 2>&1
tail -3 gpurun_out/ncu_full_$R.log
ls -la gpurun_out
