#!/usr/bin/env python
"""One pass of a scenario on the GPU (no warm-up), for use under ncu. Prints cumulative KGTable launch counts."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from harness import Session, load_product  # noqa: E402
from mceio import SHIFT_EXPLICIT, read_scenario  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "leo7"
max_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10**9
sc = read_scenario(os.path.join(ROOT, "tests", "golden", name + ".mces"))
s = Session(load_product(), sc)
cum = 0
for k, r in enumerate(sc.rec[:max_steps]):
    s.step(r)
    st = s.stats()
    cum += st.gtable_launches
    print("MU %d: gtable_launches %d cumulative %d ev_gtable_ms %.3f ev_step_ms %.3f launches %d" % (
        k + 1, st.gtable_launches, cum, st.ev_gtable_ms, st.ev_step_ms, st.kernel_launches), flush=True)
    if r.shift_kind == SHIFT_EXPLICIT:
        s.shift_b(r.delta, -1.0)
s.close()
