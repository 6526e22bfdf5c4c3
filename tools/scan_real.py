#!/usr/bin/env python
"""Times the exact scan of Re fz (KSumScan, with and without the parallel tile summaries) on the REAL slot list of a deep step: replays a scenario on the GPU,
exports the per-slot g values of the last replayed step and runs mce_debug_sum_scan on them.  Usage: scan_real.py [scenario] [steps]"""
import ctypes as ct
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from harness import Session, load_product  # noqa: E402
from mceio import SHIFT_EXPLICIT, read_scenario  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "leo7_w5"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 15
lib = load_product()
sc = read_scenario(os.path.join(ROOT, "tests", "golden", name + ".mces"))
s = Session(lib, sc)
dp = ct.POINTER(ct.c_double)
for k, r in enumerate(sc.rec[:steps]):
    s.step(r)
    if k + 1 < steps and r.shift_kind == SHIFT_EXPLICIT:
        s.shift_b(r.delta, -1.0)
n = lib.mce_debug_export_slots(s.h, 0, None, None)
g = np.zeros((n, 2))
lib.mce_debug_export_slots(s.h, n, g.ctypes.data_as(dp), None)
a = g[:, 0].copy()
run = np.add.accumulate(a)
e = np.frexp(run)[1]
print("%s step %d: %d slots, Re fz = %.17g, binade changes of the running sum %d, negative addends %.1f %%" % (name, steps, n, run[-1], int(np.sum(e[1:] != e[:-1])), 100 * np.mean(a < 0)), flush=True)
for rep in range(3):
    out = np.zeros(3)
    assert lib.mce_debug_sum_scan(s.h, n, g.ctypes.data_as(dp), out.ctypes.data_as(dp)) == 0
    print("  %s: %.3f ms (%.2f ns per slot), tiles %d, taken from summaries %d, restarts %d, exact %s" % (
        "no tiles" if os.environ.get("MCE_SCAN_NO_TILES") else "tiled", lib.mce_cpdf_last_ms(s.h), 1e6 * lib.mce_cpdf_last_ms(s.h) / n, (n + 8191) // 8192, int(out[2]), int(out[1]),
        out[0].tobytes() == np.float64(run[-1]).tobytes()), flush=True)
s.close()
if not os.environ.get("MCE_SCAN_NO_TILES") and "--both" in sys.argv:
    env = dict(os.environ); env["MCE_SCAN_NO_TILES"] = "1"
    subprocess.call([sys.executable, __file__, name, str(steps)], env=env)
