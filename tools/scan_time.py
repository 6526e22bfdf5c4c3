"""Times KSumScan (exact scan of one serial-order sum) through the debug hook: CUDA-event time of the kernel for n slots."""
import ctypes as ct, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from harness import Session, load_product
from mceio import read_scenario
lib = load_product()
s = Session(lib, read_scenario(os.path.join(ROOT, "tests", "golden", "lti3.mces"))); dp = ct.POINTER(ct.c_double)
rng = np.random.default_rng(1)
for n in (1 << 20, 1 << 23):
    g = np.zeros((n, 2)); g[:, 0] = rng.random(n) * np.exp(rng.uniform(-20, 0, n)) * np.where(rng.random(n) < 0.3, -0.2, 1.0); g = np.ascontiguousarray(g)
    out = np.zeros(2)
    for rep in range(3):
        t0 = time.perf_counter(); lib.mce_debug_sum_scan(s.h, n, g.ctypes.data_as(dp), out.ctypes.data_as(dp)); dt = time.perf_counter() - t0
    ref = np.add.accumulate(g[:, 0])[-1]
    print("n %d: kernel %.3f ms (%.2f ns per slot; the dependent chain needs 5.1), restarts %d, exact %s" % (n, lib.mce_cpdf_last_ms(s.h), lib.mce_cpdf_last_ms(s.h) * 1e6 / n, int(out[1]), out[0].tobytes() == np.float64(ref).tobytes()), flush=True)
s.close()
