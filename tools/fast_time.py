import sys, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from harness import Session, load_product
from mceio import read_scenario, SHIFT_EXPLICIT
lib=load_product(); sc=read_scenario('/root/repo/tests/golden/leo7.mces')
for fast in (False, True):
    s=Session(lib, sc, fast_moments=fast)
    for rep in range(3):
        row=[]
        for k,r in enumerate(sc.rec):
            s.step(r); st=s.stats(); row.append((st.ev_step_ms, st.ev_moments_ms))
            if r.shift_kind==SHIFT_EXPLICIT: s.shift_b(r.delta,-1.0)
        lib.mce_reset(s.h)
    print('fast' if fast else 'default', 'total %.2f'%sum(a for a,b in row), ' '.join('%.2f(%.2f)'%(a,b) for a,b in row))
    s.close()
