import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from harness import Session, load_product
from mceio import SHIFT_EXPLICIT, read_scenario
from math import comb
lib = load_product()
sc = read_scenario("/root/repo/tests/golden/leo7_w5.mces")
s = Session(lib, sc, phase_timing=True)
H = lambda m, d: sum(comb(m - 1, i) for i in range(d)) if m >= d else 2 ** (m - 1)
for rep in range(2):
    lib.mce_reset(s.h)
    for k, r in enumerate(sc.rec):
        s.step(r)
        st = s.stats()
        cnt = s.counts(False)
        cap = sum(int(cnt[m]) * H(m, sc.d) for m in range(1, len(cnt)))
        if rep == 1 and k >= 9:
            print("MU %2d total %.2f | tp %.2f mu %.2f mom %.2f regroup %.2f ftr %.2f gtable %.2f compact %.2f | chain %.2f | survivors %d cells %d capacity %d fill %.3f shapes %s" % (
                k + 1, st.ms_total, st.ms_tp, st.ms_mu, st.ms_moments, st.ms_regroup, st.ms_ftr, st.ms_gtable, st.ms_compact, st.ev_moments_ms, st.survivors, st.cells_survivors, cap,
                st.cells_survivors / max(cap, 1), {m: int(cnt[m]) for m in range(len(cnt)) if cnt[m]}), flush=True)
        if r.shift_kind == SHIFT_EXPLICIT:
            s.shift_b(r.delta, -1.0)
s.close()
