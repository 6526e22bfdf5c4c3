#!/bin/bash
# Timings, bench line, ncu launch list and one full ncu capture of the dominant kernel. Output -> gpurun_out/
mkdir -p gpurun_out
R=${1:-r01}
echo "=== full gpu test-suite"
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -12
echo "=== timings"
timeout 600 python tools/time_scenario.py leo7 2 2>&1 | tail -34
timeout 300 python tools/time_scenario.py lti3 2 2>&1 | tail -15
timeout 300 python tools/time_scenario.py leo7_w5 2 2>&1 | tail -18
echo "=== bench"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err; tail -2 gpurun_out/bench_$R.err; cat gpurun_out/bench_$R.json
echo "=== bench reference arm"
timeout 600 python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/bench_ref_$R.json 2>/dev/null; cat gpurun_out/bench_ref_$R.json
echo "=== ncu launch list (one cold pass of leo7)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$R.csv python tools/profile_pass.py leo7 > gpurun_out/profile_pass_$R.log 2>&1
tail -13 gpurun_out/profile_pass_$R.log
SKIP=$(python - <<PY
import re
c=[int(m.group(1)) for m in re.finditer(r"cumulative (\d+)", open("gpurun_out/profile_pass_$R.log").read())]
print(c[10]-1)
PY
)
echo "=== DRAM traffic of every KGTable launch"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:KGTable --csv --log-file gpurun_out/gtable_traffic_$R.csv python tools/profile_pass.py leo7 > /dev/null 2>&1
python tools/traffic_json.py gpurun_out/gtable_traffic_$R.csv > gpurun_out/traffic_$R.json; cat gpurun_out/traffic_$R.json
echo "=== ncu full capture of the last KGTable launch of MU 11 (skip $SKIP)"
timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:KGTable -s $SKIP -c 1 -o gpurun_out/prof_gtable_$R -f python tools/profile_pass.py leo7 11 > gpurun_out/ncu_full_$R.log 2>&1
tail -3 gpurun_out/ncu_full_$R.log
ls -la gpurun_out
