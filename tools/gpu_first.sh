#!/bin/bash
# First GPU contact: smoke, sanitizer on a small case, the GPU test-suite, per-step timings.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "=== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -15
echo "=== sanitizer (lti2, 5 steps)"
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -c "
import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
from harness import load_product, run_scenario
from mceio import read_scenario
sc = read_scenario('tests/golden/lti2.mces')
run_scenario(load_product(), sc, max_steps=5)
print('sanitizer run done')
" 2>&1 | tail -25
echo "=== pytest -m gpu"
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -60
echo "=== timings"
timeout 300 python tools/time_scenario.py lti3 2 2>&1 | tail -40
timeout 600 python tools/time_scenario.py leo7 1 2>&1 | tail -20
