#!/usr/bin/env python
"""Replays a scenario on the GPU and prints the per-step phase timings reported by mce_get_step_stats."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from harness import Session, load_product  # noqa: E402
from mceio import SHIFT_EXPLICIT, read_scenario  # noqa: E402

name = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
sc = read_scenario(os.path.join(ROOT, "tests", "golden", name + ".mces"))
lib = load_product()
for phases in (True, False):
    # phase timing synchronises after every phase; the pass without it is the honest wall time
    s = Session(lib, sc, phase_timing=phases)
    for rep in range(reps):
        t0 = time.time()
        print("== %s rep %d (phase timing %s)" % (name, rep, "on" if phases else "off"))
        print("step  parents    slots   muc_terms   groups survivors |  total     tp     mu    mom  regrp    ftr gtable compact | rounds launches  GB/s(alg) split ev_mu ev_mom ev_ftr ev_gtab")
        for k, r in enumerate(sc.rec):
            s.step(r)
            st = s.stats()
            gbs = st.bytes_step_algorithmic / max(st.ms_total, 1e-9) / 1e6
            print("%4d %8d %8d %11d %8d %9d | %6.2f %6.2f %6.2f %6.2f %6.2f %6.2f %6.2f %6.2f | %4d %6d %9.1f %5d %6.2f %6.2f %6.2f %6.2f" % (
                k + 1, st.parents, st.slots, st.terms_after_muc, st.groups, st.survivors, st.ms_total, st.ms_tp, st.ms_mu, st.ms_moments,
                st.ms_regroup, st.ms_ftr, st.ms_gtable, st.ms_compact, st.ftr_rounds_max, st.kernel_launches, gbs, st.split_groups, st.ev_mu_ms, st.ev_moments_ms, st.ev_ftr_ms, st.ev_gtable_ms), flush=True)
            if r.shift_kind == SHIFT_EXPLICIT:
                s.shift_b(r.delta, -1.0)
        print("total wall %.3f s" % (time.time() - t0))
        lib.mce_reset(s.h)
    s.close()
