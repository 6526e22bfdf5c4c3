#!/bin/bash
# One full ncu capture (with source) of the longest launch of a kernel (regex $1) during one cold pass of LEO7 up to MU $2. -> gpurun_out/
mkdir -p gpurun_out
K=${1:-KTpDce2}; UPTO=${2:-10}; R=${3:-r01}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:$K --csv --log-file gpurun_out/list_$K.csv python tools/profile_pass.py leo7 $UPTO > /dev/null 2>&1
SKIP=$(python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/list_$K.csv")) if len(r)>5 and r[0].isdigit()]
best=max(range(len(rows)), key=lambda i: float(rows[i][-1].replace(",","")))
print(best)
PY
)
echo "longest launch of $K: index $SKIP"
timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:$K -s $SKIP -c 1 -o gpurun_out/prof_${K}_$R -f python tools/profile_pass.py leo7 $UPTO > gpurun_out/ncu_${K}_$R.log 2>&1
tail -2 gpurun_out/ncu_${K}_$R.log
